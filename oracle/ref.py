"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/_ref/libqboot_ref.so (the UNMODIFIED reference + ref_capi.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may import this module.
Numbers travel as "qb words" (see oracle/ref_capi.cpp header): numpy uint32 arrays of shape (..., L + 2).
"""
from __future__ import annotations

import ctypes as C
import os
from fractions import Fraction

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libqboot_ref.so")

OPS = dict(
    add=0, sub=1, mul=2, div=3, fma=4, sqrt=5, mul_si=6, add_si=7, div_si=8, mul_q=9, div_q=10, add_q=11, set_q=12,
    pow=13, ui_pow=14, exp=15, log=16, gamma_inc=17, si_sub=18, si_div=19, sub_q=20, pi=21, set_si_2exp=22, cmp=23,
    isint=24, floor=25,
)

_lib = None


def available() -> bool:
    return os.path.exists(_LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_LIB_PATH)
        _lib.ref_time_gblock.restype = C.c_double
    return _lib


def nlimbs(prec: int) -> int:
    return (prec + 31) // 32


def words(prec: int) -> int:
    return nlimbs(prec) + 2


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


# ---- conversions between qb words and exact Python values -----------------------------------------------------------
def to_fraction(w, prec: int) -> Fraction:
    """Exact value of one number (zero/regular only)."""
    w = np.asarray(w, dtype=np.uint32)
    kind = int(w[0]) & 3
    if kind == 0:
        return Fraction(0)
    assert kind == 1, "inf/nan"
    L = nlimbs(prec)
    m = 0
    for i in range(L - 1, -1, -1):
        m = (m << 32) | int(w[2 + i])
    e = int(np.int32(w[1])) - 32 * L
    v = Fraction(m) * (Fraction(2) ** e)
    return -v if (int(w[0]) >> 31) else v


def from_int_exp(sign: int, m: int, e: int, prec: int) -> np.ndarray:
    """value = (-1)^sign * m * 2^e, m must fit in prec bits (exact)."""
    L = nlimbs(prec)
    w = np.zeros(L + 2, dtype=np.uint32)
    if m == 0:
        w[0] = 0x80000000 if sign else 0
        return w
    nb = m.bit_length()
    assert nb <= prec or (m >> (nb - prec)) << (nb - prec) == m, "not representable"
    sh = 32 * L - nb
    if sh >= 0:
        mm = m << sh
    else:
        mm = m >> (-sh)
    for i in range(L):
        w[2 + i] = (mm >> (32 * i)) & 0xFFFFFFFF
    w[1] = np.uint32((e + nb) & 0xFFFFFFFF)
    w[0] = 1 | (0x80000000 if sign else 0)
    return w


def from_str(s: str, prec: int) -> np.ndarray:
    out = np.zeros(words(prec), dtype=np.uint32)
    rc = lib().ref_from_str(prec, s.encode(), _p(out))
    if rc != 0:
        raise ValueError(s)
    return out


def from_rational(s: str, prec: int) -> np.ndarray:
    out = np.zeros(words(prec), dtype=np.uint32)
    rc = lib().ref_from_rational(prec, s.encode(), _p(out))
    if rc != 0:
        raise ValueError(s)
    return out


def to_str(w, prec: int, digits: int = 0) -> str:
    buf = C.create_string_buffer(4096)
    n = lib().ref_to_str(prec, _p(_u32(w)), digits, buf, 4096)
    assert n >= 0
    return buf.value.decode()


# ---- primitives ------------------------------------------------------------------------------------------------------
def op(name: str, prec: int, a=None, b=None, c=None, ia=None, ib=None, n: int | None = None) -> np.ndarray:
    W = words(prec)
    arrs = [None if x is None else _u32(x).reshape(-1, W) for x in (a, b, c)]
    ints = [None if x is None else np.ascontiguousarray(x, dtype=np.int64).reshape(-1) for x in (ia, ib)]
    if n is None:
        n = next(len(x) for x in arrs + ints if x is not None)
    out = np.zeros((n, W), dtype=np.uint32)
    rc = lib().ref_op(prec, OPS[name], n, _p(arrs[0]), _p(arrs[1]), _p(arrs[2]), _p(ints[0]), _p(ints[1]), _p(out))
    assert rc == 0
    return out


# ---- hot path --------------------------------------------------------------------------------------------------------
def dim_mixed(lam: int) -> int:
    return (lam + 2) * (lam + 2) // 4 if lam % 2 == 0 else (lam + 1) * (lam + 3) // 4


def context_consts(prec, n_max, lam, dim="3"):
    W = words(prec)
    rho = np.zeros(W, dtype=np.uint32)
    mat = np.zeros(((lam + 1), (lam + 1), W), dtype=np.uint32)
    lib().ref_context_consts(prec, n_max, lam, dim.encode(), _p(rho), _p(mat))
    return rho, mat


def rec_coeffs(prec, dim, delta, spin, S, P):
    W = words(prec)
    K, deg = (6, 3) if spin == 0 else (8, 4)
    out = np.zeros((K, deg + 1, W), dtype=np.uint32)
    k = lib().ref_rec_coeffs(prec, dim.encode(), _p(_u32(delta)), spin, _p(_u32(S)), _p(_u32(P)), _p(out))
    assert k == K
    return out


def hblock(prec, n_max, dim, delta, spin, S, P):
    W = words(prec)
    out = np.zeros((n_max + 1, W), dtype=np.uint32)
    lib().ref_hblock(prec, n_max, dim.encode(), _p(_u32(delta)), spin, _p(_u32(S)), _p(_u32(P)), _p(out))
    return out


def gblock_real(prec, n_max, lam, dim, delta, spin, S, P):
    W = words(prec)
    out = np.zeros((lam + 1, W), dtype=np.uint32)
    lib().ref_gblock_real(prec, n_max, lam, dim.encode(), _p(_u32(delta)), spin, _p(_u32(S)), _p(_u32(P)), _p(out))
    return out


def gblock(prec, n_max, lam, dim, delta, spin, S, P):
    """delta,S,P: (n, W) arrays; spin: (n,) -> (n, dim_M, W)"""
    W = words(prec)
    delta = _u32(delta).reshape(-1, W)
    n = len(delta)
    S = _u32(S).reshape(n, W)
    P = _u32(P).reshape(n, W)
    spin = np.ascontiguousarray(spin, dtype=np.uint32).reshape(n)
    dm = dim_mixed(lam)
    out = np.zeros((n, dm, W), dtype=np.uint32)
    r = lib().ref_gblock(prec, n_max, lam, dim.encode(), n, _p(delta), _p(spin), _p(S), _p(P), _p(out))
    assert r == dm, (r, dm)
    return out


def fblock(prec, n_max, lam, dim, d, delta, spin, S, P, sym):
    W = words(prec)
    out = np.zeros((dim_mixed(lam), W), dtype=np.uint32)
    n = lib().ref_fblock(prec, n_max, lam, dim.encode(), _p(_u32(d)), _p(_u32(delta)), spin, _p(_u32(S)),
                         _p(_u32(P)), int(sym), _p(out))
    return out[:n]


def conformal_scale(prec, n_max, lam, dim, spin, num_poles, include_odd, lb, ub=None):
    W = words(prec)
    cap = 128
    xs = np.zeros((cap, W), dtype=np.uint32)
    sc = np.zeros((cap, W), dtype=np.uint32)
    de = np.zeros((cap, W), dtype=np.uint32)
    bases = np.zeros((cap * cap, W), dtype=np.uint32)
    D = lib().ref_conformal_scale(prec, n_max, lam, dim.encode(), spin, num_poles, int(include_odd), _p(_u32(lb)),
                                  None if ub is None else _p(_u32(ub)), _p(xs), _p(sc), _p(de), _p(bases), cap * cap)
    assert D >= 0
    nb = D // 2 + 1
    return dict(D=D, xs=xs[: D + 1], scalings=sc[: D + 1], deltas=de[: D + 1],
                bases=bases[: nb * nb].reshape(nb, nb, W))


def build_pmp(prec, n_max, lam, dim="3", model=0, mode=0, finite=True, ds="0.5181489", de="1.412625", de1="3.83",
              numax=8, maxspin=24, threads=1, outdir=None):
    times = (C.c_double * 3)()
    counts = (C.c_int64 * 3)()
    rc = lib().ref_build_pmp(prec, n_max, lam, dim.encode(), model, mode, int(finite), ds.encode(), de.encode(),
                             de1.encode(), numax, maxspin, threads, (outdir or "").encode(), times, counts)
    assert rc == 0
    return dict(convert=times[0], create_input=times[1], write=times[2], tables=counts[0], constraints=counts[1],
                N=counts[2])


def time_gblock(prec, n_max, lam, dim, delta, spin, S, P, threads):
    W = words(prec)
    delta = _u32(delta).reshape(-1, W)
    n = len(delta)
    spin = np.ascontiguousarray(spin, dtype=np.uint32).reshape(n)
    return lib().ref_time_gblock(prec, n_max, lam, dim.encode(), n, _p(delta), _p(spin), _p(_u32(S)), _p(_u32(P)),
                                 threads)
