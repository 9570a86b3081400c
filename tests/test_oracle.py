"""The oracle pinned: the unmodified reference (oracle/_ref) against known answers, and the Python restatement
(oracle/qboot_oracle.py) against golden vectors generated from the reference (tests/golden/make_golden.py)."""
from fractions import Fraction

import numpy as np
import pytest

from oracle import qboot_oracle as qo
from tests.util import GOLDEN_OPS, golden

# Known answers computed by the reference and cross-checked against the d = 4 closed form / mpmath during the survey
# (SURVEY.md section 8(c)).  Context(400, 19, 3), 1024-bit, S = P = 0.
KAT_D3 = [
    (0, ("3.3", 4, 7), {(0, 0): "0.2416981758382605899535330879224330359716", (1, 1): "-4.324399153016532010249374883039912974892"}),
    (2, ("5.3", 4, 7), {(0, 0): "0.07600018506674183791569814688278283290538", (1, 1): "7.908059924877061975910141281992745610598"}),
]
# Context(300, 6, dim = 4), 512-bit, Delta = 5.3: equals (-1)^l 2^-l times the Dolan-Osborn closed form
KAT_D4 = [
    (0, {(0, 0): "0.14383576589682670434170090548104099216893847048431", (1, 0): "2.2025409699190164008300889311001148062738797118017",
         (2, 0): "15.897419373720925870946571454409405974383194826389", (0, 1): "-1.0410472619761054319896462583171523350372439108781"}),
    (2, {(0, 0): "0.10771351628283765246800208393154722142300897523633", (1, 0): "1.6484743565884129385263559176525414840964373655348",
         (0, 1): "0.37358552106528906346913294062262689631176082351808"}),
    (3, {(0, 0): "-0.071596149961495047216291625070930639818102388672147", (1, 0): "-1.0945137352954305274447919398666805009147556490511",
         (0, 1): "-0.92206881022685946985091752419347187550986744280183"}),
]


def _close(rec, prec, dec: str, digits=38):
    from oracle import ref as r

    v = r.to_fraction(rec, prec)
    w = Fraction(dec)
    return abs(v - w) <= abs(w) * Fraction(10) ** (-digits)


def test_reference_known_answers_d3(ref):
    prec, lam = 1024, 19
    zero = np.zeros(ref.words(prec), dtype=np.uint32)
    for spin, (base, p, q), want in KAT_D3:
        delta = ref.op("add_q", prec, a=ref.from_str(base, prec), ia=[p], ib=[q])[0]
        g = ref.gblock(prec, 400, lam, "3", delta, [spin], zero, zero)[0]
        for (m, n), dec in want.items():
            assert _close(g[qo.cf_index(lam, m, n)], prec, dec), (spin, m, n)


def test_reference_known_answers_d4(ref):
    prec, lam = 512, 6
    zero = np.zeros(ref.words(prec), dtype=np.uint32)
    delta = ref.from_str("5.3", prec)
    for spin, want in KAT_D4:
        g = ref.gblock(prec, 300, lam, "4", delta, [spin], zero, zero)[0]
        for (m, n), dec in want.items():
            assert _close(g[qo.cf_index(lam, m, n)], prec, dec, digits=40), (spin, m, n)


def test_reference_matches_committed_golden(ref):
    """the golden files are reproducible outputs of the reference"""
    g = golden("tables_small")
    prec, n_max, lam = int(g["prec"]), int(g["n_max"]), int(g["lam"])
    out = ref.gblock(prec, n_max, lam, str(g["dim"]), g["delta"], g["spin"], g["S"], g["P"])
    assert np.array_equal(out, g["g"])


@pytest.mark.parametrize("prec", [600, 1024])
def test_python_oracle_primitives_vs_golden(prec):
    g = golden(f"ops_p{prec}")
    fn = dict(add=qo.add, sub=qo.sub, mul=qo.mul, div=qo.div)
    n = len(g["a"])
    for i in range(n):
        a, b, c = (qo.from_words(g[k][i], prec) for k in "abc")
        ia, ib, ia32 = int(g["ia"][i]), int(g["ib"][i]), int(g["ia32"][i])
        for name in ("add", "sub", "mul"):
            assert np.array_equal(qo.to_words(fn[name](a, b, prec), prec), g["r_" + name][i]), (name, i)
        if not b.is_zero():
            assert np.array_equal(qo.to_words(qo.div(a, b, prec), prec), g["r_div"][i]), ("div", i)
        assert np.array_equal(qo.to_words(qo.fma(a, b, c, prec), prec), g["r_fma"][i]), ("fma", i)
        assert np.array_equal(qo.to_words(qo.mul_si(a, ia, prec), prec), g["r_mul_si"][i]), ("mul_si", i)
        assert np.array_equal(qo.to_words(qo.add_si(a, ia, prec), prec), g["r_add_si"][i]), ("add_si", i)
        assert np.array_equal(qo.to_words(qo.si_sub(ia, a, prec), prec), g["r_si_sub"][i]), ("si_sub", i)
        if ia32:
            assert np.array_equal(qo.to_words(qo.div_si(a, ia32, prec), prec), g["r_div_si"][i]), ("div_si", i)
            assert np.array_equal(qo.to_words(qo.div_q(a, Fraction(ia32, ib), prec), prec), g["r_div_q"][i]), ("div_q", i)
        q = Fraction(ia, ib)
        assert np.array_equal(qo.to_words(qo.mul_q(a, q, prec), prec), g["r_mul_q"][i]), ("mul_q", i)
        assert np.array_equal(qo.to_words(qo.add_q(a, q, prec), prec), g["r_add_q"][i]), ("add_q", i)
        assert np.array_equal(qo.to_words(qo.sub_q(a, q, prec), prec), g["r_sub_q"][i]), ("sub_q", i)
        assert np.array_equal(qo.to_words(qo.from_frac(q, prec), prec), g["r_set_q"][i]), ("set_q", i)
        pa = qo.R(False, a.m, a.e)
        assert np.array_equal(qo.to_words(qo.sqrt(pa, prec), prec), g["r_sqrt"][i]), ("sqrt", i)
    rho, _ = qo.context_consts(3, prec)
    for i in range(0, n, 4):
        y = qo.from_words(g["ys"][i], prec)
        assert np.array_equal(qo.to_words(qo.pow_(rho, y, prec), prec), g["r_pow_rho"][i]), ("pow_rho", i)
        assert np.array_equal(qo.to_words(qo.pow_(qo.from_int(4, prec), y, prec), prec), g["r_pow_4"][i]), ("pow_4", i)
    assert len(GOLDEN_OPS) == 14


@pytest.mark.parametrize("name", ["tables_small", "tables_fracdim"])
def test_python_oracle_tables_vs_golden(name):
    g = golden(name)
    prec, n_max, lam = int(g["prec"]), int(g["n_max"]), int(g["lam"])
    eps = Fraction(str(g["dim"])) / 2 - 1
    rho, mat = qo.context_consts(lam, prec)
    assert np.array_equal(qo.to_words(rho, prec), g["rho"])
    for i in range(lam + 1):
        for j in range(lam + 1):
            assert np.array_equal(qo.to_words(mat[i][j], prec), g["mat"][i, j]), (i, j)
    for t in range(len(g["spin"])):
        D, S, Pp = (qo.from_words(g[k][t], prec) for k in ("delta", "S", "P"))
        spin = int(g["spin"][t])
        b = qo.hblock(D, spin, S, Pp, eps, n_max, prec)
        assert np.array_equal(np.stack([qo.to_words(x, prec) for x in b]), g["hblock"][t]), ("hblock", t)
        if not D.is_zero():
            ps = qo.rec_coeffs(D, spin, S, Pp, eps, prec)
            for i, p in enumerate(ps):
                for d, cf in enumerate(p):
                    assert np.array_equal(qo.to_words(cf, prec), g["rec"][t, i, d]), ("rec", t, i, d)
        tab = qo.gblock(D, spin, S, Pp, eps, n_max, lam, prec)
        assert np.array_equal(np.stack([qo.to_words(x, prec) for x in tab]), g["g"][t]), ("gblock", t)
        for sym in (0, 1):
            fb = qo.fblock(qo.from_words(g["fb_d"][sym], prec), tab, sym, lam, prec)
            assert np.array_equal(np.stack([qo.to_words(x, prec) for x in fb]), g[f"fb_sym{sym}"][t]), ("fblock", t, sym)


def test_python_oracle_vs_reference_live(ref):
    """fresh seeded inputs, not in the golden files"""
    import random

    rng = random.Random(99)
    prec, n_max, lam, dim = 320, 30, 6, "3"
    eps = Fraction(1, 2)
    zero = np.zeros(ref.words(prec), dtype=np.uint32)
    for _ in range(4):
        spin = rng.choice([0, 1, 2, 5])
        delta = ref.from_str(repr(rng.uniform(0.7, 20)), prec)
        S = ref.from_str(repr(rng.uniform(-1, 1)), prec)
        g = ref.gblock(prec, n_max, lam, dim, delta, [spin], S, zero)[0]
        tab = qo.gblock(qo.from_words(delta, prec), spin, qo.from_words(S, prec), qo.ZERO, eps, n_max, lam, prec)
        assert np.array_equal(np.stack([qo.to_words(x, prec) for x in tab]), g)
