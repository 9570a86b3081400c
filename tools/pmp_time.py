"""Time a full PMP build (convert -> create_input -> write) on the GPU engine and, optionally, on the reference.
usage: python tools/pmp_time.py [single|mixed] [--ref] [--prec 1024 --nmax 400 --lam 19 --numax 20 --maxspin 50] [--repeat 2]"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qboot_b200 as qb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("model", nargs="?", default="single")
ap.add_argument("--ref", action="store_true")
ap.add_argument("--check", action="store_true", help="compare the two directories byte for byte")
ap.add_argument("--prec", type=int, default=1024)
ap.add_argument("--nmax", type=int, default=400)
ap.add_argument("--lam", type=int, default=19)
ap.add_argument("--numax", type=int, default=20)
ap.add_argument("--maxspin", type=int, default=50)
ap.add_argument("--repeat", type=int, default=2)
ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
a = ap.parse_args()
model = 0 if a.model == "single" else 1
root = tempfile.mkdtemp(prefix="qb_pmp_")
out = {"model": a.model, "threads": a.threads}
for rep in range(a.repeat):
    d = os.path.join(root, f"ours{rep}")
    t = time.time()
    r = qb.build_pmp(a.prec, a.nmax, a.lam, "3", model, 1 if model else 0, True, numax=a.numax, maxspin=a.maxspin, threads=a.threads, outdir=d)
    r["wall"] = time.time() - t
    r["bytes"] = sum(os.path.getsize(os.path.join(d, f)) for f in os.listdir(d))
    out[f"ours{rep}"] = r
if a.ref:
    from oracle import ref

    d = os.path.join(root, "ref")
    t = time.time()
    r = ref.build_pmp(a.prec, a.nmax, a.lam, "3", model, 1 if model else 0, True, numax=a.numax, maxspin=a.maxspin, threads=a.threads, outdir=d)
    r["wall"] = time.time() - t
    out["ref"] = r
    if a.check:
        import filecmp

        mine = os.path.join(root, "ours0")
        files = sorted(os.listdir(d))
        out["files"] = len(files)
        out["differing"] = [f for f in files if not filecmp.cmp(os.path.join(d, f), os.path.join(mine, f), shallow=False)]
print(json.dumps(out))
shutil.rmtree(root, ignore_errors=True)
