"""TEST INFRASTRUCTURE ONLY -- CPU restatement of qboot's block-table hot path in plain Python integers.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product never does.

What is restated (reference v0.8.1, paths relative to /root/reference):
  * the arithmetic the reference delegates to MPFR 4.2.1 (third-party, not vendored): every operation returns the exact
    result rounded once to nearest-even at `prec` bits (MPFR_RNDN, src/mp/real.cpp:5-6).  Class `R` below does that with
    exact Python integers; pow/exp go through mpmath at prec + 256 bits and are rounded once.
  * _get_rec_coeffs (include/qboot/hor_formula.hpp:34-39, src/hor_formula.cpp) from the integer tensors of
    oracle/hor_table.json (recovered by exact interpolation, oracle/derive_hor_table.py), evaluated monomial by monomial
    in lexicographic order -- the order that reproduces the reference bit for bit;
  * hBlock_shifted incl. unit and divergent operators (src/hor_recursion.cpp:14-39, include/qboot/primary_op.hpp:53-77);
  * gBlock_real (src/hor_recursion.cpp:40-56, include/qboot/algebra/real_function.hpp:236-264);
  * Context constants rho and the rho->z matrix (src/context.cpp:11-26,47-61, src/algebra/real_function.cpp:10-40,
    src/algebra/matrix.cpp:106-123);
  * Context::expand_off_diagonal (src/context.cpp:62-133), v_to_d (src/context.cpp:28-42), F_block
    (include/qboot/context.hpp:123-128, include/qboot/algebra/complex_function.hpp:190-206).

PINNING: this restatement is pinned against the reference itself -- tests/test_oracle.py checks it bit for bit against
oracle/_ref (the unmodified reference compiled here) and against tests/golden/*.npz generated from oracle/_ref by
tests/golden/make_golden.py, plus the closed-form anchors of SURVEY.md section 8(c).
"""
from __future__ import annotations

import json
import os
from fractions import Fraction
from functools import lru_cache

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class R:
    """value = (-1)^neg * m * 2^e with 0 <= m < 2^prec (m == 0: zero, sign kept); immutable."""

    __slots__ = ("neg", "m", "e")

    def __init__(self, neg, m, e):
        self.neg, self.m, self.e = bool(neg), m, e

    def is_zero(self):
        return self.m == 0

    def frac(self):
        v = Fraction(self.m) * Fraction(2) ** self.e
        return -v if self.neg else v

    def __repr__(self):
        return f"R({'-' if self.neg else '+'}{self.m}*2^{self.e})"


ZERO = R(False, 0, 0)


def _round(neg, m, e, prec, sticky=False):
    """round m * 2^e (+ a positive remainder below the last bit when sticky) to prec bits, nearest-even"""
    if m == 0:
        return R(False, 0, 0)
    nb = m.bit_length()
    if nb <= prec and not sticky:
        return R(neg, m, e)
    if nb <= prec:
        # value slightly above m: cannot reach a rounding boundary -> m with more bits is still m after rounding,
        # but make the mantissa full width first so that the sticky sits below the round bit
        sh = prec + 2 - nb
        m <<= sh
        e -= sh
        nb = prec + 2
    extra = nb - prec
    q, rem = m >> extra, m & ((1 << extra) - 1)
    half = 1 << (extra - 1)
    rnd = rem >= half
    st = (rem & (half - 1)) != 0 or sticky
    if rnd and (st or (q & 1)):
        q += 1
        if q.bit_length() > prec:
            q >>= 1
            extra += 1
    return R(neg, q, e + extra)


def from_int(v, prec):
    return _round(v < 0, abs(v), 0, prec)


def from_frac(x, prec):
    """mpfr_set_q / mpfr_set_str: exact rational rounded once"""
    x = Fraction(x)
    if x == 0:
        return ZERO
    n, d = abs(x.numerator), x.denominator
    sh = max(0, prec + 3 - (n.bit_length() - d.bit_length()))
    q, r = divmod(n << sh, d)
    return _round(x < 0, q, -sh, prec, sticky=r != 0)


def _signed(a):
    return -a.m if a.neg else a.m


def add(a, b, prec, sub=False):
    bneg = b.neg != sub
    if b.is_zero():
        if a.is_zero():
            return R(a.neg and bneg, 0, 0)
        return a
    if a.is_zero():
        return R(bneg, b.m, b.e)
    e = min(a.e, b.e)
    v = (_signed(a) << (a.e - e)) + ((-b.m if bneg else b.m) << (b.e - e))
    if v == 0:
        return ZERO
    return _round(v < 0, abs(v), e, prec)


def sub(a, b, prec):
    return add(a, b, prec, sub=True)


def neg(a):
    return R(not a.neg, a.m, a.e)


def mul(a, b, prec):
    if a.is_zero() or b.is_zero():
        return R(a.neg != b.neg, 0, 0)
    return _round(a.neg != b.neg, a.m * b.m, a.e + b.e, prec)


def fma(a, b, c, prec):
    if a.is_zero() or b.is_zero():
        if c.is_zero():
            return R((a.neg != b.neg) and c.neg, 0, 0)
        return c
    if c.is_zero():
        return mul(a, b, prec)
    pe = a.e + b.e
    e = min(pe, c.e)
    v = ((-1 if a.neg != b.neg else 1) * a.m * b.m << (pe - e)) + (_signed(c) << (c.e - e))
    if v == 0:
        return ZERO
    return _round(v < 0, abs(v), e, prec)


def div(a, b, prec):
    if a.is_zero():
        return R(a.neg != b.neg, 0, 0)
    sh = max(0, prec + 3 - (a.m.bit_length() - b.m.bit_length()))
    q, r = divmod(a.m << sh, b.m)
    return _round(a.neg != b.neg, q, a.e - b.e - sh, prec, sticky=r != 0)


def mul_si(a, v, prec):
    if v == 0 or a.is_zero():
        return R(a.neg != (v < 0), 0, 0)
    return _round(a.neg != (v < 0), a.m * abs(v), a.e, prec)


def add_si(a, v, prec):
    if v == 0:
        return a
    return add(a, from_int(v, prec), prec)


def si_sub(v, a, prec):
    if v == 0:
        return neg(a)
    if v < 0:
        return neg(add_si(a, -v, prec))
    return add(from_int(v, prec), a, prec, sub=True)


def div_si(a, v, prec):
    if a.is_zero():
        return R(a.neg != (v < 0), 0, 0)
    return div(a, R(v < 0, abs(v), 0), prec)


# rational operands: exact Fractions rounded once (mpfr_mul_q, mpfr_div_q, mpfr_add_q)
def mul_q(a, q, prec):
    q = Fraction(q)
    if a.is_zero() or q == 0:
        return R(a.neg != (q < 0), 0, 0)
    return from_frac(a.frac() * q, prec)


def div_q(a, q, prec):
    q = Fraction(q)
    if a.is_zero():
        return R(a.neg != (q < 0), 0, 0)
    return from_frac(a.frac() / q, prec)


def add_q(a, q, prec):
    q = Fraction(q)
    if q == 0:
        return a
    if a.is_zero():
        return from_frac(q, prec)
    return from_frac(a.frac() + q, prec)


def sub_q(a, q, prec):
    return add_q(a, -Fraction(q), prec)


def mul_2si(a, k):
    return a if a.is_zero() else R(a.neg, a.m, a.e + k)


def sqrt(a, prec):
    import math

    if a.is_zero():
        return a
    m, e = a.m, a.e
    sh = 2 * prec + 4 - m.bit_length()
    if (e - sh) % 2:
        sh += 1
    x = m << sh
    r = math.isqrt(x)
    return _round(False, r, (e - sh) // 2, prec, sticky=r * r != x)


def _mp_to_R(v, prec):
    from mpmath import mpf

    v = mpf(v)
    s, man, ex, _ = v._mpf_
    return _round(bool(s), int(man), int(ex), prec, sticky=True) if man else ZERO


def pow_(base: R, y: R, prec):
    """mpfr_pow / mpfr_ui_pow: correctly rounded base^y (exact cases handled by the wide intermediate)"""
    import mpmath

    if y.is_zero():
        return from_int(1, prec)
    with mpmath.workprec(prec + 256):
        b = mpmath.mpf(base.frac().numerator) / mpmath.mpf(base.frac().denominator)
        yy = mpmath.mpf(y.frac().numerator) / mpmath.mpf(y.frac().denominator)
        v = mpmath.power(b, yy)
        s, man, ex, _ = v._mpf_
    # exact results (powers of two etc.) come out with few mantissa bits: no sticky then
    exact = int(man).bit_length() <= prec
    return _round(bool(s), int(man), int(ex), prec, sticky=not exact)


# ---- record <-> R ------------------------------------------------------------------------------------------------------
def nlimbs(prec):
    return (prec + 31) // 32


def to_words(a: R, prec) -> np.ndarray:
    L = nlimbs(prec)
    w = np.zeros(L + 2, dtype=np.uint32)
    if a.is_zero():
        w[0] = 0x80000000 if a.neg else 0
        return w
    nb = a.m.bit_length()
    mm = a.m << (32 * L - nb)
    for i in range(L):
        w[2 + i] = (mm >> (32 * i)) & 0xFFFFFFFF
    w[1] = np.uint32((a.e + nb) & 0xFFFFFFFF)
    w[0] = 1 | (0x80000000 if a.neg else 0)
    return w


def from_words(w, prec) -> R:
    w = np.asarray(w, dtype=np.uint32)
    if int(w[0]) & 3 == 0:
        return R(bool(int(w[0]) >> 31), 0, 0)
    L = nlimbs(prec)
    m = 0
    for i in range(L - 1, -1, -1):
        m = (m << 32) | int(w[2 + i])
    e = int(np.int32(w[1])) - 32 * L
    tz = (m & -m).bit_length() - 1
    return R(bool(int(w[0]) >> 31), m >> tz, e + tz)


# ---- recurrence coefficients -----------------------------------------------------------------------------------------------
@lru_cache(maxsize=1)
def _hor_table():
    with open(os.path.join(_HERE, "hor_table.json")) as f:
        return json.load(f)


def _monomial(c, ex, signed, D, S, P, spin, eps, prec):
    if not signed:
        c = abs(c)
    nD, nS, nP, nl, ne = ex
    if nD:
        t = D if c == 1 else mul_si(D, c, prec)
        nD -= 1
    elif nS:
        t = S if c == 1 else mul_si(S, c, prec)
        nS -= 1
    elif nP:
        t = P if c == 1 else mul_si(P, c, prec)
        nP -= 1
    elif nl:
        t = from_int(c * spin**nl, prec)
        nl = 0
    else:
        t = from_frac(c * eps, prec)
        ne -= 1
    for _ in range(nD):
        t = mul(t, D, prec)
    for _ in range(nS):
        t = mul(t, S, prec)
    for _ in range(nP):
        t = mul(t, P, prec)
    for _ in range(nl):
        t = mul_si(t, spin, prec)
    for _ in range(ne):
        t = mul_q(t, eps, prec)
    return t


def _mul_linear(c, a, prec):
    """Polynomial::_mul_linear (src/algebra/polynomial.cpp:76-83)"""
    m = len(c) - 1
    out = [mul(c[0], a, prec)]
    for i in range(1, m + 1):
        out.append(fma(c[i], a, c[i - 1], prec))
    out.append(c[m])
    return out


def rec_coeffs(D: R, spin: int, S: R, P: R, eps: Fraction, prec):
    fam = _hor_table()["0" if spin == 0 else "1"]
    K, deg = fam["K"], fam["deg"]
    ps = [None] * K
    for i in range(1, K - 1):
        co = []
        for d in range(deg + 1):
            p = fam["polys"][f"{i},{d}"]
            mons = p["mons"]
            if not mons:
                acc = from_int(p["const"], prec)
            else:
                acc = _monomial(mons[0][0], mons[0][1], True, D, S, P, spin, eps, prec)
                for c, ex in mons[1:]:
                    t = _monomial(c, ex, False, D, S, P, spin, eps, prec)
                    acc = add(acc, t, prec, sub=c < 0)
                acc = add_si(acc, p["const"], prec)
            if p["flag"] == 1:
                lin = add(mul_si(D, -2, prec), from_frac(2 * eps, prec), prec)
                lin = add_si(lin, 3, prec)
                acc = mul(lin, acc, prec)
            co.append(acc)
        ps[i] = co
    one = from_int(1, prec)
    if spin == 0:
        p0 = [ZERO, one]
        p0 = _mul_linear(p0, add_si(D, -1, prec), prec)
        p0 = _mul_linear(p0, sub_q(mul_si(D, 2, prec), 2 * eps + 2, prec), prec)
        pl = [add_si(D, -4, prec), one]
        pl = _mul_linear(pl, add_si(mul_si(D, 2, prec), -5, prec), prec)
        pl = _mul_linear(pl, from_frac(2 * eps - 3, prec), prec)
    else:
        p0 = [ZERO, from_int(-1, prec)]
        p0 = _mul_linear(p0, add_si(D, spin - 1, prec), prec)
        p0 = _mul_linear(p0, add_si(sub_q(mul_si(D, 2, prec), 2 * eps, prec), -2, prec), prec)
        p0 = _mul_linear(p0, add_si(add_si(sub_q(D, 2 * eps, prec), -spin, prec), -1, prec), prec)
        pl = [add_si(mul_si(D, 2, prec), -7, prec), one]
        pl = _mul_linear(pl, from_frac(2 * eps - 5, prec), prec)
        pl = _mul_linear(pl, add_si(D, -(spin + 6), prec), prec)
        pl = _mul_linear(pl, add_q(D, 2 * eps + spin - 6, prec), prec)
    ps[0], ps[K - 1] = p0, pl
    return ps


def _poly_ui(c, n, prec):
    s = c[-1]
    for d in range(len(c) - 2, -1, -1):
        s = add(mul_si(s, n, prec), c[d], prec)
    return s


def leading_coefficient(spin, eps: Fraction, prec):
    num, den = Fraction(1), Fraction(1)
    for i in range(spin):
        num *= 2 * eps + i
        den *= 2 * (eps + i)
    v = num / den
    return from_frac(-v if spin % 2 else v, prec)


def _pos_int(x: R):
    return (not x.is_zero()) and (not x.neg) and x.frac().denominator == 1


def is_divergent(D: R, spin, eps, prec):
    if _pos_int(si_sub(1 - spin, D, prec)):
        return True
    if _pos_int(neg(sub_q(mul_si(D, 2, prec), 2 + 2 * eps, prec))):
        return True
    return spin > 0 and _pos_int(neg(sub_q(D, 1 + spin + 2 * eps, prec)))


def hblock(D: R, spin, S: R, P: R, eps: Fraction, n_max, prec):
    b = [leading_coefficient(spin, eps, prec)] + [ZERO] * n_max
    if D.is_zero():
        return b
    if is_divergent(D, spin, eps, prec):
        small = R(False, 1, -((3 * prec) // 8) + 15)  # C++: -(3 * prec) / 8 + 15, division truncating toward zero
        bp = hblock(mul(D, add_si(small, 1, prec), prec), spin, S, P, eps, n_max, prec)
        bm = hblock(mul(D, add_si(neg(small), 1, prec), prec), spin, S, P, eps, n_max, prec)
        return [div_si(add(x, y, prec), 2, prec) for x, y in zip(bp, bm)]
    ps = rec_coeffs(D, spin, S, P, eps, prec)
    K = len(ps)
    for n in range(1, n_max + 1):
        s = ZERO
        for i in range(1, K):
            if i <= n:
                s = fma(_poly_ui(ps[i], n, prec), b[n - i], s, prec)
        b[n] = div(neg(s), _poly_ui(ps[0], n, prec), prec)
    return b


# ---- Context constants -----------------------------------------------------------------------------------------------------
def _series_mul(x, y, prec):
    lam = len(x) - 1
    z = [ZERO] * (lam + 1)
    for k1 in range(lam + 1):
        for k2 in range(lam + 1 - k1):
            z[k1 + k2] = add(z[k1 + k2], mul(x[k1], y[k2], prec), prec)
    return z


def _converter(func, prec):
    lam = len(func) - 1
    mat = [[ZERO] * (lam + 1) for _ in range(lam + 1)]
    mat[0][0] = from_int(1, prec)
    pf = list(func)
    for n in range(1, lam + 1):
        for i in range(n, lam + 1):
            mat[n][i] = pf[i]
        pf = _series_mul(pf, func, prec)
    return mat


@lru_cache(maxsize=8)
def context_consts(lam, prec):
    sqrt8 = sqrt(from_int(8, prec), prec)
    rho = si_sub(3, sqrt8, prec)
    a = si_sub(4, sqrt8, prec)
    f = [from_frac(1 / (a.frac() ** 2), prec)]
    tmp = div(from_int(1, prec), a, prec)
    for k in range(1, lam + 1):
        t = mul(mul(f[k - 1], tmp, prec), add_si(from_int(-2, prec), -(k - 1), prec), prec)
        f.append(div_si(t, k, prec))
    func = []
    for i in range(lam + 1):
        g1 = mul(f[i - 1], sqrt8, prec) if i >= 1 else mul(ZERO, sqrt8, prec)
        g2 = div_si(f[i - 2], -2, prec) if i >= 2 else div_si(ZERO, -2, prec)
        func.append(add(g1, g2, prec))
    m0 = _converter(func, prec)
    n1 = lam + 1
    inv = [[m0[j][i] for j in range(n1)] for i in range(n1)]
    res = [[ZERO] * n1 for _ in range(n1)]
    one = from_int(1, prec)
    for i in range(n1):
        res[i][i] = div(one, inv[i][i], prec)
        for j in range(i):
            s = ZERO
            for k in range(j, i):
                s = add(s, mul(inv[i][k], res[k][j], prec), prec)
            res[i][j] = div(neg(s), inv[i][i], prec)
    fin = [ZERO] + [res[k][1] for k in range(1, n1)]
    return rho, _converter(fin, prec)


# ---- table ---------------------------------------------------------------------------------------------------------------------
def cf_index(lam, dx, dy):
    return (lam + 2 - dy) * dy + dx


def gblock_real(D, spin, S, P, eps, n_max, lam, prec):
    rho, mat = context_consts(lam, prec)
    b = hblock(D, spin, S, P, eps, n_max, prec)
    c4 = pow_(from_int(4, prec), D, prec)
    a = [mul(x, c4, prec) for x in b]
    q = D
    kf = from_int(1, prec)
    fr = []
    for k in range(lam + 1):
        if k >= 1:
            kf = mul_si(kf, k, prec)
            a = [mul(a[n], add_si(q, n, prec), prec) for n in range(n_max + 1)]
            q = add_si(q, -1, prec)
        s = ZERO
        for n in range(n_max, -1, -1):
            s = fma(s, rho, a[n], prec)
        s = mul(s, pow_(rho, q, prec), prec)
        fr.append(div(s, kf, prec) if k >= 1 else s)
    out = []
    for i in range(lam + 1):
        z = mul(fr[0], mat[0][i], prec)
        for j in range(1, lam + 1):
            z = add(z, mul(fr[j], mat[j][i], prec), prec)
        out.append(z)
    return out


def gblock(D, spin, S, P, eps, n_max, lam, prec):
    eps = Fraction(eps)
    f = [ZERO] * ((lam + 2) * (lam + 2) // 4)
    row0 = gblock_real(D, spin, S, P, eps, n_max, lam, prec)
    for m in range(lam + 1):
        f[m] = row0[m]
    t = add_si(sub_q(div_si(D, 2, prec), eps, prec), -1, prec)
    t = mul(t, D, prec)
    qc4 = mul_si(add_q(t, (2 * eps + spin) * spin / 2, prec), 4, prec)
    F = lambda mm, nn: f[cf_index(lam, mm, nn)]  # noqa: E731
    for n in range(1, lam // 2 + 1):
        common = 4 * n * eps + (4 * n - 2) * n
        for m in range(lam - 2 * n + 1):
            val = mul_si(F(m + 2, n - 1), -(m + 1) * (m + 2), prec)
            term = add_si(mul_si(add_q(S, eps, prec), 2, prec), -(m + 4 * n - 6), prec)
            term = mul(mul_si(term, 2 * (m + 1), prec), F(m + 1, n - 1), prec)
            val = add(val, term, prec)
            term = add_si(from_frac(eps * (4 * (m + n - 1)), prec), m * (m + 8 * n - 5) + n * (4 * n - 2) - 2, prec)
            term = add(term, mul_si(P, 2, prec), prec)
            term = add(term, mul_si(S, 8 * n + 4 * m - 8, prec), prec)
            term = add(term, qc4, prec)
            term = mul(mul_si(term, 4, prec), F(m, n - 1), prec)
            val = add(val, term, prec)
            if m >= 1:
                term = add_si(from_frac(eps * (2 * (m + 1 - 2 * n)), prec), m * (m + 12 * n - 13) + n * (12 * n - 34) + 22, prec)
                term = add(term, mul_si(P, 2, prec), prec)
                term = add(term, mul_si(S, 8 * n + 2 * m - 10, prec), prec)
                term = mul(mul_si(term, 8, prec), F(m - 1, n - 1), prec)
                val = add(val, term, prec)
            if n >= 2:
                term = add_si(from_frac(-2 * eps, prec), 3 * m + 4 * n - 6, prec)
                term = add(term, mul_si(S, 2, prec), prec)
                term = mul(mul_si(term, 8 * (m + 1), prec), F(m + 1, n - 2), prec)
                term = add(term, mul_si(F(m + 2, n - 2), 4 * (m + 1) * (m + 2), prec), prec)
                val = add(val, term, prec)
            val = div_q(val, common, prec)
            term = ZERO
            if m >= 3:
                term = add(term, mul_si(F(m - 3, n), 8, prec), prec)
            if m >= 2:
                term = add(term, mul_si(F(m - 2, n), 4, prec), prec)
            if m >= 1:
                term = sub(term, mul_si(F(m - 1, n), 2, prec), prec)
            f[cf_index(lam, m, n)] = add(val, term, prec)
    return f


def v_to_d(d: R, lam, prec):
    f = [ZERO] * ((lam + 2) * (lam + 2) // 4)
    f[0] = pow_(from_int(4, prec), neg(d), prec)
    md = neg(d)
    for n in range(lam // 2 + 1):
        if n > 0:
            t = mul(mul_si(f[cf_index(lam, 0, n - 1)], 4, prec), add_si(md, n - 1, prec), prec)
            f[cf_index(lam, 0, n)] = div_si(t, n, prec)
        for m in range(1, lam - 2 * n + 1):
            u = add_si(mul_si(d, -4, prec), 2 * (m + 2 * n - 1), prec)
            f[cf_index(lam, m, n)] = div_si(mul(f[cf_index(lam, m - 1, n)], u, prec), m, prec)
    return f


def fblock(d: R, g, sym, lam, prec):
    v = v_to_d(d, lam, prec)
    out = []
    for N in range(lam // 2 + 1):
        for M in range(lam - 2 * N + 1):
            if M % 2 != sym:
                continue
            acc = ZERO
            for dy1 in range(N + 1):
                for dx1 in range(M + 1):
                    acc = add(acc, mul(v[cf_index(lam, dx1, dy1)], g[cf_index(lam, M - dx1, N - dy1)], prec), prec)
            out.append(mul_si(acc, 2, prec))
    return out
