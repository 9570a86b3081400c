"""Host side of the C++ façade (include/qboot, qboot_b200/host) against the reference run live (oracle/_ref): the
arithmetic behind qboot::mp::real -- including the correctly rounded transcendentals that replace MPFR's -- decimal
input / output, and ConformalScale.  No GPU involved: none of this is on the device path."""
import ctypes as C
import random
from fractions import Fraction

import numpy as np
import pytest

from tests.util import P, same_records, words


def hop(lib, ref, name, prec, a=None, b=None, c=None, ia=None, ib=None, n=None):
    W = words(prec)
    arrs = [None if x is None else np.ascontiguousarray(x, dtype=np.uint32).reshape(-1, W) for x in (a, b, c)]
    ints = [None if x is None else np.ascontiguousarray(x, dtype=np.int64).reshape(-1) for x in (ia, ib)]
    if n is None:
        n = next(len(x) for x in arrs + ints if x is not None)
    out = np.zeros((n, W), dtype=np.uint32)
    rc = lib.qbh_hostop(prec, ref.OPS.get(name, 26), n, P(arrs[0]), P(arrs[1]), P(arrs[2]), P(ints[0]), P(ints[1]), P(out))
    assert rc == 0, (name, rc)
    return out


@pytest.mark.parametrize("prec", [512, 600, 800, 1000, 1024, 1200, 1536])
def test_real_ops_vs_mpfr(hostfacade, ref, prec):
    from tests.golden.make_golden import rnd_num

    rng = random.Random(prec)
    n = 40
    a = np.stack([rnd_num(rng, prec) for _ in range(n)])
    b = np.stack([rnd_num(rng, prec) for _ in range(n)])
    c = np.stack([rnd_num(rng, prec) for _ in range(n)])
    ia = np.array([rng.choice([0, 1, -1, 2, -3, 7, 400, -12345, 2**31 - 1, 2**40 + 3, rng.randint(-10**6, 10**6)]) for _ in range(n)], dtype=np.int64)
    ib = np.array([rng.choice([1, 2, 3, 7, 10, 20, 1000, 2**32 - 1, 10**10, rng.randint(1, 10**6)]) for _ in range(n)], dtype=np.int64)
    pos = a.copy()
    pos[:, 0] &= 0x7FFFFFFF
    small = np.stack([rnd_num(rng, prec, -6, 6) for _ in range(n)])
    cases = [("add", dict(a=a, b=b)), ("sub", dict(a=a, b=b)), ("mul", dict(a=a, b=b)), ("div", dict(a=a, b=b)),
             ("fma", dict(a=a, b=b, c=c)), ("sqrt", dict(a=pos)), ("mul_si", dict(a=a, ia=ia)), ("add_si", dict(a=a, ia=ia)),
             ("div_si", dict(a=a, ia=ia)), ("si_sub", dict(a=a, ia=ia)), ("si_div", dict(a=a, ia=ia)),
             ("mul_q", dict(a=a, ia=ia, ib=ib)), ("div_q", dict(a=a, ia=ia, ib=ib)), ("add_q", dict(a=a, ia=ia, ib=ib)),
             ("sub_q", dict(a=a, ia=ia, ib=ib)), ("set_q", dict(ia=ia, ib=ib)), ("exp", dict(a=small)), ("log", dict(a=pos)),
             ("pow", dict(a=pos, b=small)), ("pi", dict(n=1)), ("floor", dict(a=a)), ("isint", dict(a=a)), ("cmp", dict(a=a, b=b))]
    for name, kw in cases:
        assert same_records(hop(hostfacade, ref, name, prec, **kw), ref.op(name, prec, **kw)), name
    # Gamma(0, x) over the range a scale factor asks for (1e-10 pole shifts up to spin 80 poles)
    xs = np.stack([ref.from_str(s, prec) for s in ["3.76e-11", "1e-5", "0.01", "0.37", "1", "2.5", "7.25", "19.3", "33.1", "75.5", "120.25"]])
    z = np.zeros_like(xs)
    assert same_records(hop(hostfacade, ref, "gamma_inc", prec, a=z, b=xs), ref.op("gamma_inc", prec, a=z, b=xs))
    # integer powers: exact range and the exp/log range
    for nn in [2, 3, -2, -5, -41, 60, -60]:
        bb = np.stack([ref.from_int_exp(int(nn < 0), abs(nn), 0, prec)] * n)
        assert same_records(hop(hostfacade, ref, "pow_si", prec, a=a, ia=np.full(n, nn, dtype=np.int64)), ref.op("pow", prec, a=a, b=bb)), nn
    # powers with exactly representable results
    xa = np.stack([ref.from_str(s, prec) for s in ["4", "4", "2.25", "0.25", "9", "1", "16", "0.6862915010152396"]])
    xb = np.stack([ref.from_str(s, prec) for s in ["1.5", "-0.5", "0.5", "3.5", "0.5", "7.7", "0.25", "3"]])
    assert same_records(hop(hostfacade, ref, "pow", prec, a=xa, b=xb), ref.op("pow", prec, a=xa, b=xb))


@pytest.mark.parametrize("prec", [600, 1024])
def test_decimal_text_vs_mpfr(hostfacade, ref, prec):
    from tests.golden.make_golden import rnd_num

    for s in ["1.409", "1e-10", "0.5181489", "3.83", "-0.125", "6", "1.4135", "123456789012345678901234567890.5e-7", "-7.25e33", ".5", "5."]:
        out = np.zeros(words(prec), dtype=np.uint32)
        assert hostfacade.qbh_from_str(prec, s.encode(), P(out)) == 0
        assert np.array_equal(out, ref.from_str(s, prec)), s
    rng = random.Random(7 * prec)
    vals = [rnd_num(rng, prec, -900, 900) for _ in range(60)]
    vals += [ref.from_str(s, prec) for s in ["1", "0.5", "10", "1000", "999999", "1e-5", "0.001", "0.0001", "1e300", "1e-300", "123.456", "0.1", "-2"]]
    # a tie at the SDPB digit count: 0.5 + 2^-(nd + 1) needs round-half-even
    nd = 3 + int(prec * 0.302)
    vals.append(ref.op("add", prec, a=ref.from_str("0.5", prec), b=ref.from_int_exp(0, 1, -(nd + 1), prec))[0])
    buf = C.create_string_buffer(8192)
    for v in vals:
        for digits in (0, 6, 1, 20):
            k = hostfacade.qbh_to_str(prec, P(np.ascontiguousarray(v)), digits, buf, 8192)
            assert k >= 0
            assert buf.value.decode() == ref.to_str(v, prec, digits), (ref.to_str(v, prec, 20), digits)


CASES = [(0, 8, False, "0.5", None), (0, 8, False, "1.409", "1.4135"), (0, 8, True, "6", None), (2, 9, False, "5", None),
         (4, 10, False, None, None), (1, 8, True, None, None), (3, 9, True, "5.25", "5.33"), (24, 12, False, None, None),
         (7, 11, True, None, None), (0, 0, False, "3", None), (50, 30, False, None, None)]


@pytest.mark.parametrize("prec,n_max,lam,dim,eps", [(600, 100, 11, "3", (1, 2)), (800, 200, 13, "3", (1, 2)), (1024, 400, 19, "3", (1, 2)),
                                                     (1000, 100, 10, "37/10", (17, 20)), (1536, 100, 27, "4", (1, 1))])
def test_conformal_scale_vs_reference(hostfacade, ref, prec, n_max, lam, dim, eps):
    """poles, sample points, sample Deltas (incl. the Moebius map of finite ranges), sample scalings and the bilinear
    bases (E_1 moments + Hankel-Cholesky) -- all bit-identical to the reference's ConformalScale"""
    W = words(prec)
    cap = 128
    for spin, npoles, odd, lb, ub in CASES:
        if lb is None:
            e = Fraction(*eps)
            g = e if spin == 0 else spin + 2 * e
            lbw = ref.from_rational(f"{g.numerator}/{g.denominator}", prec)
        else:
            lbw = ref.from_str(lb, prec)
        ubw = None if ub is None else ref.from_str(ub, prec)
        xs, sc, de = (np.zeros((cap, W), dtype=np.uint32) for _ in range(3))
        ba = np.zeros((cap * cap, W), dtype=np.uint32)
        D = hostfacade.qbh_conformal_scale(prec, lam, C.c_int64(eps[0]), C.c_int64(eps[1]), spin, npoles, int(odd), P(lbw),
                                           None if ubw is None else P(ubw), P(xs), P(sc), P(de), P(ba), cap * cap)
        want = ref.conformal_scale(prec, n_max, lam, dim, spin, npoles, odd, lbw, ubw)
        assert D == want["D"]
        nb = D // 2 + 1
        assert np.array_equal(xs[: D + 1], want["xs"])
        assert np.array_equal(de[: D + 1], want["deltas"])
        assert np.array_equal(sc[: D + 1], want["scalings"])
        assert np.array_equal(ba[: nb * nb].reshape(nb, nb, W), want["bases"]), (spin, npoles, odd, lb, ub)
