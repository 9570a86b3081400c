"""Build the CUDA engine in-tree: qboot_b200/libqboot_b200.so (sm_100a only).

One translation unit per limb count (engine_inst.cu with -DQB_L=<L>) plus the C ABI (capi.cu), compiled in parallel
with nvcc and linked against the shared CUDA runtime.  No JIT cache, no torch extension machinery: the resulting .so
sits next to this file so that it travels with the repository snapshot to the GPU box.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
BUILD = os.path.join(ROOT, "build", "qb")
LIB = os.path.join(HERE, "libqboot_b200.so")

# limb counts compiled in: 512, 600, 800, 1000/1024, 1200, 1536 bits
LIMBS = (16, 19, 25, 32, 38, 48)

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _sources_digest() -> str:
    h = hashlib.sha256()
    for d, _, files in sorted(os.walk(CSRC)):
        for f in sorted(files):
            with open(os.path.join(d, f), "rb") as fh:
                h.update(f.encode())
                h.update(fh.read())
    with open(os.path.join(ROOT, "include", "qboot_b200.h"), "rb") as fh:
        h.update(fh.read())
    h.update(repr(LIMBS).encode())
    return h.hexdigest()


def _run(cmd, log):
    with open(log, "w") as fh:
        p = subprocess.run(cmd, stdout=fh, stderr=subprocess.STDOUT)
    if p.returncode != 0:
        sys.stderr.write(open(log).read()[-4000:])
        raise RuntimeError("build step failed: " + " ".join(cmd))


def build(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(BUILD, exist_ok=True)
    stamp = os.path.join(BUILD, "stamp")
    digest = _sources_digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    jobs = []
    for L in LIMBS:
        obj = os.path.join(BUILD, f"engine_L{L}.o")
        cmd = [NVCC, *ARCH, *COMMON, f"-DQB_L={L}", "-c", os.path.join(CSRC, "engine_inst.cu"), "-o", obj]
        jobs.append((cmd, os.path.join(BUILD, f"engine_L{L}.log"), obj))
    llist = " ".join(f"X({L})" for L in LIMBS)
    capi_obj = os.path.join(BUILD, "capi.o")
    jobs.append(([NVCC, *ARCH, *COMMON, f"-DQB_L_LIST={llist}", "-c", os.path.join(CSRC, "capi.cu"), "-o", capi_obj],
                 os.path.join(BUILD, "capi.log"), capi_obj))
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        futs = [ex.submit(_run, cmd, log) for cmd, log, _ in jobs]
        for f in futs:
            f.result()
    objs = [o for _, _, o in jobs]
    _run([NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-cudart", "shared"], os.path.join(BUILD, "link.log"))
    with open(stamp, "w") as fh:
        fh.write(digest)
    if verbose:
        print(f"built {LIB}")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
