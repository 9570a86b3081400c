"""Shared helpers of the test-suite."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

OPS = dict(add=0, sub=1, mul=2, div=3, fma=4, sqrt=5, mul_si=6, add_si=7, div_si=8, mul_q=9, div_q=10, add_q=11,
           set_q=12, pow=13, ui_pow=14, exp=15, si_sub=18, sub_q=20, cmp=23, isint=24)


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def P(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def words(prec):
    return (prec + 31) // 32 + 2


def same_records(r, e):
    """bitwise equality of records, treating inf/nan as equal when the kinds agree"""
    r = np.asarray(r)
    e = np.asarray(e)
    if r.shape != e.shape:
        return False
    flat_r = r.reshape(-1, r.shape[-1])
    flat_e = e.reshape(-1, e.shape[-1])
    for x, y in zip(flat_r, flat_e):
        if not np.array_equal(x, y) and not ((x[0] & 3) >= 2 and (x[0] & 3) == (y[0] & 3)):
            return False
    return True


def emu_op(emu, name, prec, a=None, b=None, c=None, ia=None, ib=None, consts=None):
    W = words(prec)
    arrs = [None if x is None else np.ascontiguousarray(x, dtype=np.uint32).reshape(-1, W) for x in (a, b, c)]
    ints = [None if x is None else np.ascontiguousarray(x, dtype=np.int64).reshape(-1) for x in (ia, ib)]
    n = next(len(x) for x in arrs + ints if x is not None)
    out = np.zeros((n, W), dtype=np.uint32)
    rc = emu.qbh_op(prec, OPS[name], n, P(arrs[0]), P(arrs[1]), P(arrs[2]), P(ints[0]), P(ints[1]), P(out), P(consts))
    assert rc == 0, (name, rc)
    return out


# operands of every golden primitive: (name, operand keys)
GOLDEN_OPS = [("add", ("a", "b")), ("sub", ("a", "b")), ("mul", ("a", "b")), ("div", ("a", "b")), ("fma", ("a", "b", "c")),
              ("mul_si", ("a", "ia")), ("add_si", ("a", "ia")), ("div_si", ("a", "ia32")), ("si_sub", ("a", "ia")),
              ("mul_q", ("a", "ia", "ib")), ("div_q", ("a", "ia32", "ib")), ("add_q", ("a", "ia", "ib")),
              ("sub_q", ("a", "ia", "ib")), ("set_q", ("ia", "ib"))]


def op_kwargs(g, keys):
    kw = {}
    for k in keys:
        kw["ia" if k == "ia32" else k] = g[k]
    return kw
