"""Host build of the engine's __host__ __device__ arithmetic and per-table stages (tests/hostemu) against golden vectors
from the reference and against the reference live.  This is the CPU-side guard for the CUDA code: same source, host
compiler.  (The product itself has no CPU path; these entry points exist only in the test harness.)"""
import ctypes as C
import random
from fractions import Fraction

import numpy as np
import pytest

from tests.util import GOLDEN_OPS, P, emu_op, golden, op_kwargs, same_records, words


def _ext_consts(ref, prec, rho):
    """[ln2, ln4, ln rho] at L + 3 limbs from MPFR (harness-side constants for pow)"""
    E = ref.nlimbs(prec) + 3
    pe = 32 * E
    ln2 = ref.op("log", pe, a=ref.from_int_exp(0, 2, 0, pe))[0]
    ln4 = ref.op("log", pe, a=ref.from_int_exp(0, 4, 0, pe))[0]
    fr = ref.to_fraction(rho, prec)
    rho_w = ref.from_int_exp(0, fr.numerator, -(fr.denominator.bit_length() - 1), pe)
    return np.concatenate([ln2, ln4, ref.op("log", pe, a=rho_w)[0]])


@pytest.mark.parametrize("prec", [600, 1024])
def test_primitives_vs_golden(hostemu, prec):
    g = golden(f"ops_p{prec}")
    for name, keys in GOLDEN_OPS:
        kw = op_kwargs(g, keys)
        assert same_records(g["r_" + name], emu_op(hostemu, name, prec, **kw)), name
    pos = g["a"].copy()
    pos[:, 0] &= 0x7FFFFFFF
    out = np.zeros_like(pos)
    hostemu.qbh_sqrt(prec, len(pos), P(pos), P(out))
    assert same_records(g["r_sqrt"], out)


@pytest.mark.parametrize("prec", [64, 65, 100, 128, 150, 256, 512, 800, 1000, 1200, 1536])
def test_primitives_vs_mpfr_live(hostemu, ref, prec):
    rng = random.Random(prec)
    from tests.golden.make_golden import rnd_num

    n = 120
    a = np.stack([rnd_num(rng, prec) for _ in range(n)])
    b = np.stack([rnd_num(rng, prec) for _ in range(n)])
    c = np.stack([rnd_num(rng, prec) for _ in range(n)])
    for i in range(0, n, 3):
        b[i] = a[i]
        b[i, 2] ^= (rng.getrandbits(8) << ((-prec) % 32)) & 0xFFFFFFFF
        if a[i, 0] & 3:
            b[i, 0] ^= 0x80000000
    ia = np.array([rng.choice([0, 1, -1, 2, -3, 7, 400, -12345, 2**31 - 1, 2**40 + 3, rng.randint(-10**6, 10**6)]) for _ in range(n)], dtype=np.int64)
    ib = np.array([rng.choice([1, 2, 3, 7, 10, 20, 1000, 2**32 - 1, rng.randint(1, 10**6)]) for _ in range(n)], dtype=np.int64)
    ia32 = np.clip(ia, -(2**32 - 1), 2**32 - 1)
    prod = ref.op("mul", prec, a=a, b=b)
    prod[:, 0] ^= 0x80000000
    cases = [("add", dict(a=a, b=b)), ("sub", dict(a=a, b=b)), ("mul", dict(a=a, b=b)), ("div", dict(a=a, b=b)),
             ("fma", dict(a=a, b=b, c=c)), ("fma", dict(a=a, b=b, c=prod)), ("mul_si", dict(a=a, ia=ia)),
             ("add_si", dict(a=a, ia=ia)), ("div_si", dict(a=a, ia=ia32)), ("si_sub", dict(a=a, ia=ia)),
             ("mul_q", dict(a=a, ia=ia, ib=ib)), ("div_q", dict(a=a, ia=ia32, ib=ib)), ("add_q", dict(a=a, ia=ia, ib=ib)),
             ("sub_q", dict(a=a, ia=ia, ib=ib)), ("set_q", dict(ia=ia, ib=ib))]
    for name, kw in cases:
        assert same_records(ref.op(name, prec, **kw), emu_op(hostemu, name, prec, **kw)), name
    for name, kw in [("cmp", dict(a=a, b=b)), ("isint", dict(a=a))]:
        assert np.array_equal(ref.op(name, prec, **kw)[:, 0], emu_op(hostemu, name, prec, **kw)[:, 0]), name
    pos = a.copy()
    pos[:, 0] &= 0x7FFFFFFF
    out = np.zeros_like(pos)
    hostemu.qbh_sqrt(prec, n, P(pos), P(out))
    assert same_records(ref.op("sqrt", prec, a=pos), out)


@pytest.mark.parametrize("prec", [128, 150, 256, 600, 608, 800, 1024, 1200, 1536])
def test_division_quotient_digit_corrections(hostemu, ref, prec):
    """Knuth's long division estimates every quotient digit from the top limbs and is occasionally one too large; the
    multiply-subtract then borrows and the divisor is added back -- about once in 2^32 (2^64 on the host's 64-bit limbs)
    for random operands, so random tests never see that branch.  Forced here: a divisor whose mantissa is 1000...0001 (or
    1000...0 followed by a low tail) over numerators q * 1000...0 makes the estimate q while the digit is q - 1, in the
    first digit and, through the shifted numerators, in later ones; also divisors of all ones (every estimate saturates)."""
    L = (prec + 31) // 32
    low = 1 << (32 * L - prec)  # lowest bit of a prec-bit mantissa inside its limbs
    top = 1 << (32 * L - 1)

    def rec(mant, exp=0, neg=False):
        r = np.zeros(L + 2, dtype=np.uint32)
        r[0] = 1 | (0x80000000 if neg else 0)
        r[1] = np.uint32(exp & 0xFFFFFFFF)
        for j in range(L):
            r[2 + j] = (mant >> (32 * j)) & 0xFFFFFFFF
        return r

    divisors = [top | low, top | (low << 1) | low, top | (low << 40) | low, top | ((1 << 31) * low), (1 << (32 * L)) - low,
                top | (((1 << 64) - 1) * low), top | (1 << (32 * L - 70))]
    numerators = [top, top | low, top | (top >> 1), top | (top >> 33), top | (top >> 64), top | (top >> 65) | low,
                  (1 << (32 * L)) - low, top | (low << 3)]
    a, b = [], []
    for d in divisors:
        for nmr in numerators:
            a.append(rec(nmr, 1))
            b.append(rec(d, -3, neg=True))
    a, b = np.stack(a), np.stack(b)
    want = ref.op("div", prec, a=a, b=b)
    assert same_records(want, emu_op(hostemu, "div", prec, a=a, b=b))
    # ... and through the device's register-resident division (fastops.cuh: 3-by-2 reciprocal, add-back)
    if prec in (608, 800, 1024, 1536):
        assert same_records(want, _fast(hostemu, FAST["div"], prec, a=a, b=b))
    want = ref.op("mul", prec, a=a, b=b)
    assert same_records(want, emu_op(hostemu, "mul", prec, a=a, b=b))


@pytest.mark.parametrize("prec", [600, 1024])
def test_pow_vs_golden(hostemu, ref, prec):
    g = golden(f"ops_p{prec}")
    rho = ref.op("si_sub", prec, a=ref.op("sqrt", prec, a=ref.from_int_exp(0, 8, 0, prec)), ia=[3])[0]
    ce = _ext_consts(ref, prec, rho)
    E2 = ref.nlimbs(prec) + 3 + 2
    ln2, ln4, lnrho = ce[:E2], ce[E2:2 * E2], ce[2 * E2:]
    n = len(g["ys"])
    got = emu_op(hostemu, "pow", prec, a=np.tile(rho, (n, 1)), b=g["ys"], consts=np.concatenate([ln2, lnrho]))
    assert np.array_equal(got, g["r_pow_rho"])
    got = emu_op(hostemu, "ui_pow", prec, a=g["ys"], consts=np.concatenate([ln2, ln4]))
    assert np.array_equal(got, g["r_pow_4"])


@pytest.mark.parametrize("prec,lam", [(600, 11), (1024, 19), (256, 5)])
def test_context_constants(hostemu, ref, prec, lam):
    """rho, the rho->z matrix (bit-exact) and the extended logarithms (<= 1 ulp at L + 3 limbs)"""
    W, E = words(prec), ref.nlimbs(prec) + 3
    rho, mat = ref.context_consts(prec, 30, lam, "3")
    r2 = np.zeros(W, dtype=np.uint32)
    m2 = np.zeros((lam + 1, lam + 1, W), dtype=np.uint32)
    kf = np.zeros((lam + 1, W), dtype=np.uint32)
    ext = np.zeros((3, E + 2), dtype=np.uint32)
    assert hostemu.qbh_context_consts(prec, lam, P(r2), P(m2), P(kf), P(ext)) == 0
    assert np.array_equal(rho, r2)
    assert np.array_equal(mat, m2)
    want = _ext_consts(ref, prec, rho).reshape(3, E + 2)
    pe = 32 * E
    for i in range(3):
        a, b = ref.to_fraction(want[i], pe), ref.to_fraction(ext[i], pe)
        assert abs(a - b) <= abs(a) * Fraction(2) ** (-pe + 1), i


def _run_table(hostemu, ref, g, t, ce):
    prec, n_max, lam = int(g["prec"]), int(g["n_max"]), int(g["lam"])
    W = words(prec)
    eps = Fraction(str(g["dim"])) / 2 - 1
    dm = (lam + 2) * (lam + 2) // 4
    out = np.zeros((dm, W), dtype=np.uint32)
    b0 = g["hblock"][t][0].copy()
    k = hostemu.qbh_gblock(prec, n_max, lam, C.c_int64(eps.numerator), eps.denominator, P(g["delta"][t].copy()),
                           int(g["spin"][t]), P(g["S"][t].copy()), P(g["P"][t].copy()), P(b0), P(g["rho"].copy()),
                           P(np.ascontiguousarray(g["mat"])), P(ce), P(out), None)
    assert k == dm
    return out


@pytest.mark.parametrize("name", ["tables_small", "tables_c1", "tables_fracdim"])
def test_table_stages_vs_golden(hostemu, ref, name):
    g = golden(name)
    prec, n_max, lam = int(g["prec"]), int(g["n_max"]), int(g["lam"])
    W = words(prec)
    eps = Fraction(str(g["dim"])) / 2 - 1
    ce = _ext_consts(ref, prec, g["rho"])
    for t in range(len(g["spin"])):
        spin = int(g["spin"][t])
        if (int(g["delta"][t][0]) & 3) == 0 or (spin == 2 and t == 1 and str(g["dim"]) == "3"):
            continue  # unit / divergent operators are planned by the engine's host side (covered by the GPU tests)
        K, deg = (6, 3) if spin == 0 else (8, 4)
        rec = np.zeros((K, deg + 1, W), dtype=np.uint32)
        assert hostemu.qbh_rec_coeffs(prec, C.c_int64(eps.numerator), eps.denominator, P(g["delta"][t].copy()), spin,
                                      P(g["S"][t].copy()), P(g["P"][t].copy()), P(rec)) == K
        assert np.array_equal(rec, g["rec"][t, :K, :deg + 1]), ("rec", t)
        hb = np.zeros((n_max + 1, W), dtype=np.uint32)
        hostemu.qbh_hblock(prec, n_max, C.c_int64(eps.numerator), eps.denominator, P(g["delta"][t].copy()), spin,
                           P(g["S"][t].copy()), P(g["P"][t].copy()), P(g["hblock"][t][0].copy()), P(hb))
        assert np.array_equal(hb, g["hblock"][t]), ("hblock", t)
        tab = _run_table(hostemu, ref, g, t, ce)
        assert np.array_equal(tab, g["g"][t]), ("gblock", t)
        for sym in (0, 1):
            want = g[f"fb_sym{sym}"][t]
            fo = np.zeros_like(want)
            cnt = hostemu.qbh_fblock(prec, lam, P(g["fb_d"][sym].copy()), P(tab), sym, P(ce), P(fo))
            assert cnt == len(want) and np.array_equal(fo, want), ("fblock", t, sym)


# ---- register-resident fast paths (fastops.cuh): host build with the carry chains emulated ----------------------------------
FAST = dict(mul=40, fma=41, add=42, sub=43, mul_si=44, add_si=45, div=46)


def _fast(emu, code, prec, a=None, b=None, c=None, ia=None):
    W = words(prec)
    arrs = [None if x is None else np.ascontiguousarray(x, dtype=np.uint32).reshape(-1, W) for x in (a, b, c)]
    ints = None if ia is None else np.ascontiguousarray(ia, dtype=np.int64)
    n = len(arrs[0])
    out = np.zeros((n, W), dtype=np.uint32)
    assert emu.qbh_op(prec, code, n, P(arrs[0]), P(arrs[1]), P(arrs[2]), P(ints), None, P(out), None) == 0
    return out


def fast_cases(ref, prec, n, seed):
    """random operands plus the adversarial ones: c ~ -(a*b) with single-bit perturbations (cancellation of every depth) and
    exponent gaps around 0, 32, 64, prec and beyond in both directions (rounding residues of the series are ~2^-prec)"""
    from tests.golden.make_golden import rnd_num

    rng = random.Random(seed)
    L = (prec + 31) // 32
    a = np.stack([rnd_num(rng, prec, -5, 5) for _ in range(n)])
    b = np.stack([rnd_num(rng, prec, -5, 5) for _ in range(n)])
    c = np.stack([rnd_num(rng, prec, -80, 80) for _ in range(n)])
    prod = ref.op("mul", prec, a=a, b=b)
    gaps = [0, 0, 1, -1, 31, 32, 33, 63, 64, 65, -31, -32, -33, -64, -65, prec - 33, prec, prec + 1, prec + 33, prec + 64, prec + 65, 3000,
            -prec, -prec - 1, -prec - 65, -3000]
    for i in range(0, n, 2):
        c[i] = prod[i].copy()
        if (c[i, 0] & 3) == 1:
            c[i, 0] ^= 0x80000000
            if rng.random() < 0.8:
                li = rng.randrange(L)
                c[i, 2 + li] ^= 1 << rng.randrange(31 if li == L - 1 else 32)
            c[i, 2] &= (0xFFFFFFFF << ((-prec) % 32)) & 0xFFFFFFFF
            c[i, 1] = np.uint32((int(np.int32(c[i, 1])) + rng.choice(gaps + [rng.randint(-prec - 80, prec + 80)])) & 0xFFFFFFFF)
    ia = np.array([rng.choice([1, -1, 2, 3, 7, 400, -12345, 2**31 - 1, 2**32 - 1, -(2**32 - 1)]) for _ in range(n)], dtype=np.int64)
    b2 = b.copy()
    for i in range(0, n, 3):
        b2[i] = a[i]
        li = rng.randrange(L)
        b2[i, 2 + li] ^= 1 << rng.randrange(31 if li == L - 1 else 32)
        b2[i, 2] &= (0xFFFFFFFF << ((-prec) % 32)) & 0xFFFFFFFF
    ok = [i for i in range(n) if (b2[i, 0] & 3) == 1 and (b[i, 0] & 3) == 1]
    return [("mul", dict(a=a, b=b)), ("fma", dict(a=a, b=b, c=c)), ("add", dict(a=prod, b=c)), ("sub", dict(a=prod, b=c)),
            ("add", dict(a=a, b=b)), ("mul_si", dict(a=a, ia=ia)), ("add_si", dict(a=a, ia=ia)), ("add_si", dict(a=c, ia=ia)),
            ("div", dict(a=a[ok], b=b[ok])), ("div", dict(a=a[ok], b=b2[ok]))]


@pytest.mark.parametrize("prec", [64, 96, 100, 128, 600, 800, 1000, 1024, 1200, 1536])
def test_fast_paths_vs_mpfr(hostemu, ref, prec):
    for name, kw in fast_cases(ref, prec, 400, prec):
        want = ref.op(name, prec, **kw)
        got = _fast(hostemu, FAST[name], prec, **kw)
        bad = [i for i in range(len(want)) if not np.array_equal(want[i], got[i])]
        assert not bad, (name, bad[:5])


@pytest.mark.parametrize("prec", [512, 608, 800, 1024, 1216, 1536])
def test_short_product_boundaries(hostemu, ref, prec):
    """The multiply fast path builds its L+3-limb frame from a SHORT product (fastops.cuh::mul_top_short) and must notice
    when the skipped low limb products could matter.  Operands are constructed so that the low limbs of the exact product
    are chosen at will (a = target * b^-1 modulo 2^(32 (L-3)), b odd): the two limbs just below the frame sit on and around
    every decision boundary (0, L, 2^32 - L, 2^32 - 1) over all-zero, all-one and random tails."""
    rng = random.Random(prec)
    L = prec // 32
    low_bits = 32 * (L - 3)
    mod = 1 << low_bits
    M32 = 0xFFFFFFFF
    f4s = [0, 1, 2, L - 6, L - 5, L - 4, L - 3, L, L + 1, M32 - L - 2, M32 - L - 1, M32 - L, M32 - L + 1, M32 - 4, M32 - 1, M32]
    A, B = [], []
    for f4 in f4s:
        for f5 in (0, 1, M32, rng.getrandbits(32)):
            for tail in ("zeros", "ones", "random"):
                nb = 32 * (L - 5)
                t = {"zeros": 0, "ones": (1 << nb) - 1, "random": rng.getrandbits(nb)}[tail]
                target = (f4 << (32 * (L - 4))) | (f5 << (32 * (L - 5))) | t
                b = rng.getrandbits(prec) | (1 << (prec - 1)) | 1
                a_low = (target * pow(b, -1, mod)) % mod
                a = (rng.getrandbits(96) | (1 << 95)) << low_bits | a_low
                assert (a * b) % mod == target
                A.append(ref.from_int_exp(rng.getrandbits(1), a, rng.randint(-40, 40) - prec, prec))
                B.append(ref.from_int_exp(rng.getrandbits(1), b, rng.randint(-40, 40) - prec, prec))
    # short operands (machine numbers, few-digit decimals): the limbs just below the frame are exactly zero
    for _ in range(60):
        wide = rng.getrandbits(prec) | (1 << (prec - 1)) | 1
        nbits = rng.choice([1, 2, 24, 53, 64, 96, 128, 160, 32 * 5, 32 * 5 + 1, 32 * 6])
        short = (rng.getrandbits(nbits) | (1 << (nbits - 1)) | 1) << (prec - nbits)
        x, y = (wide, short) if rng.random() < 0.7 else (short, wide)
        A.append(ref.from_int_exp(rng.getrandbits(1), x, rng.randint(-40, 40) - prec, prec))
        B.append(ref.from_int_exp(rng.getrandbits(1), y, rng.randint(-40, 40) - prec, prec))
    A, B = np.stack(A), np.stack(B)
    from tests.golden.make_golden import rnd_num

    C = np.stack([rnd_num(rng, prec, -80, 80) for _ in range(len(A))])
    for name, kw in (("mul", dict(a=A, b=B)), ("mul", dict(a=B, b=A)), ("fma", dict(a=A, b=B, c=C))):
        want = ref.op(name, prec, **kw)
        got = _fast(hostemu, FAST[name], prec, **kw)
        bad = [i for i in range(len(want)) if not np.array_equal(want[i], got[i])]
        assert not bad, (name, bad[:5])


def test_ziv_window_flags_exactly_the_values_next_to_a_midpoint(hostemu, ref):
    """pow_from_log reports a rounding as in doubt iff the extended value lies inside the window around a midpoint between
    two prec-bit numbers (mpf_math.cuh: ziv_safe).  With the window widened to 1/16 ulp on either side the flagged set
    must be exactly the set found from MPFR's value of the same power at 256 more bits."""
    import math
    import random
    from fractions import Fraction

    import qboot_b200 as qb
    from tests.util import OPS

    prec, wide = 600, 600 + 256
    rho = ref.op("si_sub", prec, a=ref.op("sqrt", prec, a=ref.from_int_exp(0, 8, 0, prec)), ia=[3])[0]
    ce = _ext_consts(ref, prec, rho)
    E2 = ref.nlimbs(prec) + 3 + 2
    ln2, lnrho = ce[:E2], ce[2 * E2:]
    guard = 32 * (ref.nlimbs(prec) + 3) - prec
    rng = random.Random(7)
    ys = [Fraction(rng.randint(1, 10**8), rng.randint(10**5, 10**6)) for _ in range(160)]
    rho_w = qb.from_fraction(qb.to_fraction(rho, prec), wide)
    flagged = truth = 0
    for y in ys:
        yr = qb.from_fraction(y, prec)
        out = np.zeros((1, words(prec)), dtype=np.uint32)
        rc = hostemu.qbh_op(prec, OPS["pow"], 1, P(rho[None, :].copy()), P(yr[None, :].copy()), None,
                            P(np.array([guard - 4], dtype=np.int64)), None, P(out), P(np.concatenate([ln2, lnrho])))
        assert rc in (0, -2)
        v = qb.to_fraction(ref.op("pow", wide, a=rho_w, b=qb.from_fraction(qb.to_fraction(yr, prec), wide))[0], wide)
        e = v.numerator.bit_length() - v.denominator.bit_length()
        while Fraction(2) ** (e - 1) > v:
            e -= 1
        while Fraction(2) ** e <= v:
            e += 1
        f = v / Fraction(2) ** (e - prec)
        f -= math.floor(f)
        # f in [7/16, 9/16): round bit 1 followed by 000, or round bit 0 followed by 111
        near = Fraction(7, 16) <= f < Fraction(9, 16)
        assert (rc == -2) == near, (y, float(f), rc)
        flagged += rc == -2
        truth += near
        # whatever the window says, the value itself is MPFR's
        assert np.array_equal(out[0], ref.op("pow", prec, a=rho, b=yr)[0])
    assert flagged == truth and 5 <= flagged <= 40


@pytest.mark.parametrize("prec", [64, 100, 600, 800, 1000, 1024, 1200, 1536])
def test_fused_horner_step_vs_mpfr(hostemu, ref, prec):
    """fast::horner_step -- s <- RN(RN(s n) + c), the step of the reference's Polynomial::eval with an integer argument -- against
    mpfr_mul_ui followed by mpfr_add, on operands built to hit its corners: products whose dropped bits are exactly a tie
    or one off a tie, all-ones mantissas (the rounding carry runs through every limb), addends that cancel the product to
    every depth, exponent gaps around 0, 32, 64 and prec in both directions, n = 1 and n close to 2^32."""
    from tests.golden.make_golden import rnd_num

    rng = random.Random(prec * 7 + 1)
    L = (prec + 31) // 32
    drop = 32 * L - prec
    S, Nn, Cc = [], [], []
    ns = [1, 2, 3, 5, 7, 8, 100, 255, 256, 399, 400, 401, 65535, 65536, 2**31 - 1, 2**31, 2**32 - 1]
    for it in range(700):
        n = rng.choice(ns) if it % 3 else rng.randint(1, 400)
        style = it % 7
        if style == 0:      # mantissa chosen so that m * n has its dropped bits at / around the tie
            k = n.bit_length()  # up to k bits are dropped
            m = (rng.getrandbits(prec) | (1 << (prec - 1)))
            target_low = rng.choice([1 << (k - 1), (1 << (k - 1)) + 1, (1 << (k - 1)) - 1, 0, (1 << k) - 1]) if k else 0
            if n % 2 == 1 and k:
                inv = pow(n, -1, 1 << k)
                m = (m & ~((1 << k) - 1)) | ((target_low * inv) % (1 << k))
            m |= 1 << (prec - 1)
        elif style == 1:    # all ones: the rounding carry propagates to the top
            m = (1 << prec) - 1
            if rng.random() < 0.5:
                m ^= rng.getrandbits(4)
        else:
            m = rng.getrandbits(prec) | (1 << (prec - 1))
        e = rng.randint(-6, 12)
        s = ref.from_int_exp(rng.getrandbits(1), m, e - prec, prec)
        S.append(s)
        Nn.append(n)
        Cc.append(rnd_num(rng, prec, -6, 24))
    S, Cc = np.stack(S), np.stack(Cc)
    Nn = np.array(Nn, dtype=np.int64)
    prod = ref.op("mul_si", prec, a=S, ia=Nn)
    gaps = [0, 0, 0, 1, -1, 2, 31, 32, 33, 63, 64, 65, -31, -32, -33, -63, -64, -65, prec - 1, prec, prec + 1, prec + 70, -prec, -prec - 1, -prec - 70]
    for i in range(0, len(S), 2):   # addends that cancel the product: -product with one bit flipped, at every exponent gap
        if (prod[i, 0] & 3) != 1:
            continue
        Cc[i] = prod[i].copy()
        Cc[i, 0] ^= 0x80000000
        if rng.random() < 0.8:
            li = rng.randrange(L)
            Cc[i, 2 + li] ^= 1 << rng.randrange(31 if li == L - 1 else 32)
        Cc[i, 2] &= (0xFFFFFFFF << drop) & 0xFFFFFFFF
        Cc[i, 1] = np.uint32((int(np.int32(Cc[i, 1])) + rng.choice(gaps)) & 0xFFFFFFFF)
    want = ref.op("add", prec, a=prod, b=Cc)
    W = words(prec)
    out = np.zeros((len(S), W), dtype=np.uint32)
    assert hostemu.qbh_op(prec, 47, len(S), P(S), P(Cc), None, P(Nn), None, P(out), None) == 0
    declined = int(((out[:, 0] & 0x100) != 0).sum())
    out[:, 0] &= ~np.uint32(0x100)
    bad = [i for i in range(len(S)) if not np.array_equal(want[i], out[i])]
    assert not bad, (bad[:5], [int(Nn[i]) for i in bad[:5]])
    assert declined < len(S) // 2   # the fast path takes the bulk (zeros, deep cancellations and the like fall back)


@pytest.mark.parametrize("prec", [64, 100, 600, 800, 1000, 1024, 1200, 1536])
def test_fixed_point_polynomials_vs_mpfr(hostemu, ref, prec):
    """csrc/polyfx.cuh -- the recurrence polynomials evaluated in a fixed-point frame with the reference's roundings emulated
    on integers -- against mpfr_mul_ui + mpfr_add Horner (Polynomial::eval, polynomial.hpp:66-83) for n = 1 .. n_lim and
    beyond, on polynomials shaped like the path's (coefficients of comparable size), with alternating signs (cancellation
    between the terms), with roots at or next to integers (deep cancellation: must fall back, not be wrong), with zero
    coefficients, tiny coefficients (not representable in the frame: fall back) and coefficients of few bits (ties)."""
    from tests.golden.make_golden import rnd_num

    rng = random.Random(prec * 11 + 3)
    W = words(prec)
    nbits = 9
    ns = np.array(list(range(1, 64)) + [100, 127, 128, 255, 256, 399, 400, 401, 500, 511], dtype=np.uint32)
    fx_used = total = 0
    for it in range(60):
        deg = rng.choice([3, 4, 4, 4, 1, 2])
        style = it % 6
        cs = []
        for d in range(deg + 1):
            if style == 0:      # comparable magnitudes, random signs
                c = rnd_num(rng, prec, -4, 30)
            elif style == 1:    # c_d ~ n_typ^-d: every term of the Horner sum has the same size (alternating signs cancel)
                m = rng.getrandbits(prec) | (1 << (prec - 1))
                c = ref.from_int_exp(d % 2, m, 20 - 8 * d - prec, prec)
            elif style == 2:    # few-bit coefficients: products and sums hit ties and exact values
                m = rng.getrandbits(rng.randint(1, 12)) | 1
                c = ref.from_int_exp(rng.getrandbits(1), m, rng.randint(-6, 6), prec)
            elif style == 3:    # one coefficient is zero, one is tiny (2^-300 below the others: outside any frame)
                c = rnd_num(rng, prec, 0, 8)
                if d == 1:
                    c = ref.from_int_exp(0, 0, 0, prec)
                if d == 0 and it % 12 == 3:
                    c = rnd_num(rng, prec, -330, -300)
            elif style == 4:    # (n - r)(...) with an integer root r: exact zeros and neighbours of zero
                c = None
            else:
                c = rnd_num(rng, prec, -40, 40)
            cs.append(c)
        if style == 4:
            r = rng.randint(1, 60)
            # p(n) = (n - r) (a n^2 + b n + g) with small integer a, b, g: integer coefficients, root at n = r
            a, b, g = rng.randint(1, 2**20), rng.randint(-2**20, 2**20), rng.randint(-2**20, 2**20)
            ints = [-r * g, g - r * b, b - r * a, a]
            deg = 3
            cs = [ref.from_int_exp(int(v < 0), abs(v), 0, prec) for v in ints]
        coefs = np.stack(cs)
        out = np.zeros((len(ns), W), dtype=np.uint32)
        status = np.zeros(len(ns), dtype=np.int32)
        assert hostemu.qbh_poly_fx(prec, deg, nbits, P(coefs), len(ns), P(ns), P(out), P(status)) == 0
        # the reference's Horner, vectorised over n
        s = np.tile(coefs[deg], (len(ns), 1))
        for d in range(deg - 1, -1, -1):
            s = ref.op("add", prec, a=ref.op("mul_si", prec, a=s, ia=ns.astype(np.int64)), b=np.tile(coefs[d], (len(ns), 1)))
        bad = [int(ns[i]) for i in range(len(ns)) if not np.array_equal(s[i], out[i])]
        assert not bad, (it, style, bad[:5])
        fx_used += int(status.sum())
        total += len(ns)
    assert fx_used > total // 2   # the fixed-point path carries the bulk; the rest fell back (and is still exact)
