"""The PMP half of the path without a GPU: the C++ facade (planner, key dedupe, PMP assembly / pack / decimal text from
the engine's host+device source, SDPB writer) against the reference's convert -> create_input -> write, file by file.
Block tables come from the oracle here (tests/hostemu/emu_capi.cpp); tests/test_gpu.py repeats the comparison with the
CUDA library on the B200."""
import filecmp
import os

import pytest

import qboot_b200 as qb


def same_dirs(a, b):
    fa, fb = sorted(os.listdir(a)), sorted(os.listdir(b))
    assert fa == fb
    return [f for f in fa if not filecmp.cmp(os.path.join(a, f), os.path.join(b, f), shallow=False)]


CASES = [
    # prec, n_max, lam, model, mode, finite, numax, maxspin
    (600, 100, 11, 0, 0, True, 8, 24),     # BASELINE config 1: single-correlator Ising, 329 tables, FindContradiction
    (600, 60, 9, 0, 1, True, 6, 12),       # ExtremalOPE on the same model
    (800, 60, 9, 1, 1, True, 6, 10),       # mixed Ising with the finite ranges [lb, ub), ExtremalOPE
    (800, 40, 7, 1, 0, False, 5, 9),       # mixed Ising, FindContradiction
    (1000, 50, 8, 0, 0, True, 5, 8),       # a precision that is not a limb multiple
]


@pytest.mark.parametrize("prec,n_max,lam,model,mode,finite,numax,maxspin", CASES)
def test_sdpb_directory_equals_reference(facade_emu, ref, tmp_path, prec, n_max, lam, model, mode, finite, numax, maxspin):
    want, got = str(tmp_path / "ref"), str(tmp_path / "ours")
    r = ref.build_pmp(prec, n_max, lam, "3", model, mode, finite, numax=numax, maxspin=maxspin, threads=4, outdir=want)
    o = qb.build_pmp(prec, n_max, lam, "3", model, mode, finite, numax=numax, maxspin=maxspin, threads=4, outdir=got, lib=facade_emu)
    assert o["constraints"] == r["constraints"] and o["N"] == r["N"]
    assert same_dirs(got, want) == []          # byte-identical files: blocks, bilinear bases, objectives, d_B, d_c


def test_thread_count_does_not_change_the_output(facade_emu, tmp_path):
    a, b = str(tmp_path / "p1"), str(tmp_path / "p8")
    qb.build_pmp(600, 40, 7, "3", 1, 1, True, numax=4, maxspin=6, threads=1, outdir=a, lib=facade_emu)
    qb.build_pmp(600, 40, 7, "3", 1, 1, True, numax=4, maxspin=6, threads=8, outdir=b, lib=facade_emu)
    assert same_dirs(a, b) == []


def test_fractional_dimension(facade_emu, ref, tmp_path):
    want, got = str(tmp_path / "ref"), str(tmp_path / "ours")
    ref.build_pmp(600, 40, 7, "37/10", 0, 0, False, ds="0.8", de="2.2", numax=4, maxspin=6, threads=2, outdir=want)
    qb.build_pmp(600, 40, 7, "37/10", 0, 0, False, ds="0.8", de="2.2", numax=4, maxspin=6, threads=2, outdir=got, lib=facade_emu)
    assert same_dirs(got, want) == []


REF_SRC = "/root/reference"


@pytest.mark.parametrize("src,theirs,args", [
    ("sample/main.cpp", "sample_ref", ()),            # single-correlator Ising, 1000-bit, n_Max 400, lambda 16, FindContradiction
    ("test/src/main.cpp", "testmain_ref", ("solve",)),  # mixed Ising, 1000-bit, n_Max 400, lambda 14, ExtremalOPE
])
def test_reference_drivers_compile_unmodified_against_the_facade(facade_emu, ref, tmp_path, src, theirs, args):
    """The reference's own driver programs, UNMODIFIED, compiled against include/qboot and linked with the facade (here over
    the host stand-in of the C ABI; tests/test_gpu.py runs the copies linked with the CUDA library): same SDPB directory,
    byte for byte, as the same source linked with the reference."""
    import subprocess

    from tests.conftest import ROOT

    theirs = os.path.join(ROOT, "oracle", "_ref", theirs)
    if not (os.path.isdir(REF_SRC) and os.path.exists(theirs)):
        pytest.skip("needs /root/reference (driver sources) and oracle/_ref")
    bdir = os.path.join(ROOT, "tests", "_build")
    exe = str(tmp_path / "driver_emu")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-DNDEBUG", "-w", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(REF_SRC, src), "-o", exe, "-L" + bdir, "-lqboot_facade_emu", "-lqb_emu",
                           "-Wl,-rpath," + bdir, "-pthread"])
    dirs = []
    for prog, cwd in ((exe, tmp_path / "ours"), (theirs, tmp_path / "ref")):
        os.makedirs(cwd)
        subprocess.run([prog, *args], cwd=cwd, check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900)
        (d,) = os.listdir(cwd)
        dirs.append((d, str(cwd / d)))
    assert dirs[0][0] == dirs[1][0]
    assert same_dirs(dirs[0][1], dirs[1][1]) == []
