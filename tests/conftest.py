import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def hostemu():
    """Host build of the engine's __host__ __device__ arithmetic (test harness only, tests/hostemu)."""
    import ctypes

    out = os.path.join(ROOT, "tests", "_build", "libqb_hostemu.so")
    srcs = [os.path.join(ROOT, "tests", "hostemu", f) for f in ("hostemu_ops.cpp", "hostemu_block.cpp")]
    deps = srcs + [os.path.join(ROOT, "qboot_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "qboot_b200", "csrc"))]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-o", out, *srcs])
    return ctypes.CDLL(out)


def _build_host_lib(out, srcs, deps, extra=()):
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-w", "-I" + os.path.join(ROOT, "include"), "-o", out, *srcs, *extra])


def _tree(*dirs):
    out = []
    for d in dirs:
        for base, _, files in os.walk(os.path.join(ROOT, d)):
            out += [os.path.join(base, f) for f in files if f.endswith((".cpp", ".hpp", ".h", ".cuh", ".inc"))]
    return out


@pytest.fixture(scope="session")
def hostfacade():
    """Host arithmetic + ConformalScale of the C++ façade behind test-only C entry points (tests/hostemu/*_capi.cpp)."""
    import ctypes

    out = os.path.join(ROOT, "tests", "_build", "libqb_hostops_test.so")
    srcs = [os.path.join(ROOT, "tests", "hostemu", f) for f in ("hostops_capi.cpp", "facade_scale_capi.cpp")]
    srcs += [os.path.join(ROOT, "qboot_b200", "host", f) for f in ("hostops.cpp", "facade_math.cpp")]
    _build_host_lib(out, srcs, srcs + _tree("include", "qboot_b200/host", "qboot_b200/csrc"))
    return ctypes.CDLL(out)


@pytest.fixture(scope="session")
def facade_emu(ref):
    """The C++ facade linked against a host stand-in of the C ABI (tests/hostemu/emu_capi.cpp: tables from the oracle,
    PMP stages from the engine's own host+device source).  Lets the CPU suite run convert -> create_input -> write."""
    import qboot_b200 as qb

    bdir = os.path.join(ROOT, "tests", "_build")
    emu = os.path.join(bdir, "libqb_emu.so")
    deps = _tree("include", "qboot_b200/host", "qboot_b200/csrc", "tests/hostemu")
    _build_host_lib(emu, [os.path.join(ROOT, "tests", "hostemu", "emu_capi.cpp"), os.path.join(ROOT, "qboot_b200", "host", "hostops.cpp")],
                    deps, ["-ldl"])
    fac = os.path.join(bdir, "libqboot_facade_emu.so")
    srcs = [os.path.join(ROOT, "qboot_b200", "host", f) for f in ("facade_math.cpp", "facade_context.cpp", "facade_program.cpp", "models_capi.cpp")]
    _build_host_lib(fac, srcs, deps + [emu], ["-L" + bdir, "-lqb_emu", "-Wl,-rpath,$ORIGIN", "-pthread"])
    os.environ["QB_EMU_REF"] = os.path.join(ROOT, "oracle", "_ref", "libqboot_ref.so")
    return qb.load_facade(fac)


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference (oracle/_ref); built on demand where /root/reference exists."""
    from oracle import ref as r

    if not r.available():
        if os.path.isdir("/root/reference/src"):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref", "-j8"])
        else:
            pytest.skip("oracle/_ref not built and /root/reference absent")
    return r
