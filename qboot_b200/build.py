"""Build the CUDA engine in-tree: qboot_b200/libqboot_b200.so (sm_100a only).

Translation units: per limb count L one host-side engine (engine_inst.cu, -DQB_L) and seven kernel groups
(kernel_inst.cu, -DQB_L -DQB_KGROUP=0..6), plus the C ABI (capi.cu).  All device arithmetic is force-inlined into the
kernels (see mpf.cuh), so the kernel groups are the expensive part; they compile in parallel and are cached per object
by a digest of its own preprocessed source.  No JIT cache, no torch extension machinery: the .so sits next to this file so that it
travels with the repository snapshot to the GPU box.
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
# QB_BUILD_DIR / QB_LIB_OUT: development builds of a variant next to the in-tree library (the facade is then not relinked)
BUILD = os.environ.get("QB_BUILD_DIR") or os.path.join(ROOT, "build", "qb")
LIB = os.environ.get("QB_LIB_OUT") or os.path.join(HERE, "libqboot_b200.so")
# the C++ facade (include/qboot/*.hpp + qboot_b200/host/*.cpp): plain host C++17, reaches CUDA only through the C ABI above
FACADE = os.path.join(HERE, "libqboot.so")
HOST = os.path.join(HERE, "host")
HOST_SRCS = ("hostops.cpp", "facade_math.cpp", "facade_context.cpp", "facade_program.cpp", "models_capi.cpp")
# the system g++ on purpose: $CXX in this image names a toolchain that links its own libstdc++ statically, and two
# libstdc++ copies in one process (this library + the nvcc-built one) crash in iostream code
CXX = os.environ.get("QB_CXX", "/usr/bin/g++")

# limb counts compiled in: 512, 600, 800, 1000/1024, 1200, 1536 bits (override: QB_LIMBS=19,32)
LIMBS = tuple(int(x) for x in os.environ.get("QB_LIMBS", "16,19,25,32,38,48").split(","))
KGROUPS = (0, 1, 2, 3, 4, 5, 6, 7)

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _tu_digest(cmd) -> str:
    """digest of ONE translation unit: its preprocessed source (so only the headers it includes matter) + the flags"""
    i = cmd.index("-c")
    pre = subprocess.run([cmd[0], *[a for a in cmd[1:i] if a not in ("-lineinfo",)], "-E", cmd[i + 1]],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
    h = hashlib.sha256()
    # drop line markers: they carry absolute paths
    h.update(b"\n".join(l for l in pre.stdout.split(b"\n") if not l.startswith(b"#")))
    h.update(" ".join(cmd[:i]).encode())
    return h.hexdigest()


def _run(cmd, log, stamp=None, digest=None):
    with open(log, "w") as fh:
        p = subprocess.run(cmd, stdout=fh, stderr=subprocess.STDOUT)
    if p.returncode != 0:
        sys.stderr.write(open(log).read()[-4000:])
        raise RuntimeError("build step failed: " + " ".join(cmd))
    if stamp:
        with open(stamp, "w") as fh:
            fh.write(digest)


def build(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(BUILD, exist_ok=True)
    jobs = []  # (cmd, log, obj)

    def add(name, src, defs):
        obj = os.path.join(BUILD, name + ".o")
        jobs.append(([NVCC, *ARCH, *COMMON, *defs, "-c", os.path.join(CSRC, src), "-o", obj], os.path.join(BUILD, name + ".log"), obj))

    for L in LIMBS:
        for g in KGROUPS:
            # QB_NVCC_EXTRA_G<g>: extra flags for one kernel group (development: try a launch configuration without
            # rebuilding the other groups)
            extra = os.environ.get(f"QB_NVCC_EXTRA_G{g}", "").split()
            add(f"kernels_L{L}_g{g}", "kernel_inst.cu", [f"-DQB_L={L}", f"-DQB_KGROUP={g}", f"-DQB_L_FIRST={LIMBS[0]}", *extra])
    for L in LIMBS:
        add(f"engine_L{L}", "engine_inst.cu", [f"-DQB_L={L}"])
    add("capi", "capi.cu", ["-DQB_L_LIST=" + " ".join(f"X({L})" for L in LIMBS)])

    todo = []
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        keys = list(ex.map(lambda j: _tu_digest(j[0]), jobs))
    for (cmd, log, obj), key in zip(jobs, keys):
        stamp = obj + ".stamp"
        if force or not (os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == key):
            todo.append((cmd, log, stamp, key))
    if todo or not os.path.exists(LIB):
        # heaviest first (series / deriv of the largest L)
        todo.sort(key=lambda t: (-int(next((a.split("=")[1] for a in t[0] if a.startswith("-DQB_L=")), "0")),))
        with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
            futs = [ex.submit(_run, cmd, log, stamp, key) for cmd, log, stamp, key in todo]
            for f in futs:
                f.result()
        objs = [o for _, _, o in jobs]
        _run([NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-cudart", "shared"], os.path.join(BUILD, "link.log"))
        if verbose:
            print(f"built {LIB} ({len(todo)} objects compiled)")
    if not os.environ.get("QB_LIB_OUT"):
        build_facade(force=force, verbose=verbose)
    return LIB


def build_facade(force: bool = False, verbose: bool = True, engine_lib: str = LIB, out: str = FACADE) -> str:
    """g++ build of the C++ facade into qboot_b200/libqboot.so, linked against the C-ABI library next to it."""
    os.makedirs(BUILD, exist_ok=True)
    inc = os.path.join(ROOT, "include")
    deps = [os.path.join(HOST, f) for f in os.listdir(HOST)]
    for base, _, files in os.walk(inc):
        deps += [os.path.join(base, f) for f in files]
    deps += [os.path.join(CSRC, f) for f in ("mpf.cuh", "fastops.cuh", "mpf_math.cuh", "hostmath.hpp")]
    tag = os.path.basename(out)
    objs, todo = [], []
    for src in HOST_SRCS:
        obj = os.path.join(BUILD, f"{tag}.{src}.o")
        objs.append(obj)
        if force or not os.path.exists(obj) or any(os.path.getmtime(d) > os.path.getmtime(obj) for d in deps):
            todo.append(([CXX, "-std=c++17", "-O2", "-g", "-fPIC", "-pthread", "-w", "-I" + inc, "-c", os.path.join(HOST, src), "-o", obj],
                         os.path.join(BUILD, f"{tag}.{src}.log")))
    if todo or not os.path.exists(out) or os.path.getmtime(engine_lib) > os.path.getmtime(out):
        with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
            for f in [ex.submit(_run, cmd, log) for cmd, log in todo]:
                f.result()
        libdir, libname = os.path.split(engine_lib)
        _run([CXX, "-shared", "-o", out, *objs, "-L" + libdir, "-l:" + libname, "-Wl,-rpath,$ORIGIN", "-pthread"],
             os.path.join(BUILD, f"{tag}.link.log"))
        if verbose:
            print(f"built {out} ({len(todo)} objects compiled)")
    return out


if __name__ == "__main__":
    build(force="--force" in sys.argv)
