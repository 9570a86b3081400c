"""The C-ABI library: loads, exports every symbol include/qboot_b200.h declares, and refuses to run without a GPU."""
import ctypes as C
import os
import re

import pytest

import qboot_b200 as qb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "qboot_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qb_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(qb.SYMBOLS)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(qb.LIB_PATH):
        pytest.skip("library not built yet (python -m qboot_b200.build)")
    lib = C.CDLL(qb.LIB_PATH)
    for name in _declared_symbols():
        assert hasattr(lib, name), name


def test_no_cpu_fallback():
    """without a CUDA device the engine must fail loudly, never compute on the host"""
    if not os.path.exists(qb.LIB_PATH):
        pytest.skip("library not built yet")
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(qb.QbError):
        qb.Engine(1024)
    lib = qb.load_library()
    assert lib.qb_precision_supported(1024) == 1 and lib.qb_precision_supported(4096) == 0
    assert b"no CPU fallback" in lib.qb_strerror(-1)


def test_number_conversions_roundtrip():
    from fractions import Fraction

    for prec in (64, 600, 1024):
        for s in ("0.5181489", "-3", "1/3", "1.412625e-3", "12345678901234567890.5"):
            w = qb.from_str(s, prec)
            v = qb.to_fraction(w, prec)
            assert abs(v - Fraction(s)) <= abs(Fraction(s)) * Fraction(2) ** (-prec)
        assert qb.to_fraction(qb.from_fraction(0, prec), prec) == 0


def test_number_conversion_matches_mpfr(ref):
    for prec in (100, 600, 1024):
        for s in ("0.5181489", "-3", "1.412625", "3.83", "12345678901234567890.5", "1e-30"):
            assert (qb.from_str(s, prec) == ref.from_str(s, prec)).all(), (prec, s)
