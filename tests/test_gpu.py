"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against golden vectors generated from the
unmodified reference and against the reference itself (oracle/_ref, shipped as a prebuilt .so) on seeded inputs.

Bar: bit-exact.  Every operation of the engine is the reference's operation (exact result rounded once to nearest-even
at `prec` bits, in the reference's order), so tables must equal the MPFR build's tables word for word -- far inside the
north-star tolerance of 2^-(prec-32) relative.
"""
import random
from fractions import Fraction

import numpy as np
import pytest

import qboot_b200 as qb
from tests.util import GOLDEN_OPS, OPS, golden, op_kwargs, same_records

pytestmark = pytest.mark.gpu


def _engine(prec, n_max, lam, dim="3"):
    return qb.Engine(prec).context(n_max, lam, dim)


@pytest.mark.parametrize("prec", [600, 1024])
def test_device_primitives_vs_golden(prec):
    g = golden(f"ops_p{prec}")
    eng = _engine(prec, 20, 5)
    for name, keys in GOLDEN_OPS:
        got = eng.op(OPS[name], **op_kwargs(g, keys))
        assert same_records(g["r_" + name], got), name
    n = len(g["ys"])
    assert np.array_equal(eng.op(OPS["pow"], b=g["ys"]), g["r_pow_rho"])
    assert np.array_equal(eng.op(OPS["ui_pow"], a=g["ys"]), g["r_pow_4"])
    assert n > 0
    eng.close()


@pytest.mark.parametrize("prec", [512, 800, 1200, 1536])
def test_device_primitives_vs_mpfr_live(ref, prec):
    from tests.golden.make_golden import rnd_num

    rng = random.Random(prec)
    n = 200
    a = np.stack([rnd_num(rng, prec) for _ in range(n)])
    b = np.stack([rnd_num(rng, prec) for _ in range(n)])
    c = np.stack([rnd_num(rng, prec) for _ in range(n)])
    for i in range(0, n, 3):
        b[i] = a[i]
        b[i, 2] ^= (rng.getrandbits(8) << ((-prec) % 32)) & 0xFFFFFFFF
        if a[i, 0] & 3:
            b[i, 0] ^= 0x80000000
    ia = np.array([rng.choice([0, 1, -1, 2, -3, 7, 400, -12345, 2**31 - 1, 2**40 + 3]) for _ in range(n)], dtype=np.int64)
    ib = np.array([rng.choice([1, 2, 3, 7, 10, 20, 1000, 2**32 - 1]) for _ in range(n)], dtype=np.int64)
    ia32 = np.clip(ia, -(2**32 - 1), 2**32 - 1)
    eng = _engine(prec, 20, 5)
    for name, kw in [("add", dict(a=a, b=b)), ("sub", dict(a=a, b=b)), ("mul", dict(a=a, b=b)), ("div", dict(a=a, b=b)),
                     ("fma", dict(a=a, b=b, c=c)), ("mul_si", dict(a=a, ia=ia)), ("add_si", dict(a=a, ia=ia)),
                     ("div_si", dict(a=a, ia=ia32)), ("si_sub", dict(a=a, ia=ia)), ("mul_q", dict(a=a, ia=ia, ib=ib)),
                     ("div_q", dict(a=a, ia=ia32, ib=ib)), ("add_q", dict(a=a, ia=ia, ib=ib)), ("set_q", dict(ia=ia, ib=ib))]:
        assert same_records(ref.op(name, prec, **kw), eng.op(OPS[name], **kw)), name
    eng.close()


@pytest.mark.parametrize("name", ["tables_c1", "tables_fracdim", "tables_c3"])
def test_tables_vs_golden(name):
    """gBlock tables incl. the unit operator and the divergent stress tensor, Context constants and F/H blocks"""
    g = golden(name)
    prec, n_max, lam, dim = int(g["prec"]), int(g["n_max"]), int(g["lam"]), str(g["dim"])
    eng = _engine(prec, n_max, lam, dim)
    rho, mat = eng.context_consts()
    assert np.array_equal(rho, g["rho"]) and np.array_equal(mat, g["mat"])
    got = eng.gblock(g["spin"], g["delta"], g["S"], g["P"])
    assert np.array_equal(got, g["g"])
    n = len(g["spin"])
    for sym in (0, 1):
        fb = eng.fblock(list(range(n)), g["fb_d"][sym:sym + 1], [0] * n, [sym] * n)
        assert np.array_equal(np.stack(fb), g[f"fb_sym{sym}"]), sym
    eng.close()


def _random_inputs(rng, prec, n, spins=(0, 1, 2, 3, 4, 8, 17, 30, 51)):
    spin = np.array([rng.choice(spins) for _ in range(n)], dtype=np.uint32)
    delta = np.stack([qb.from_str(repr(rng.uniform(0.55, 60.0)), prec) for _ in range(n)])
    zero = qb.from_fraction(0, prec)
    S = np.stack([qb.from_str(repr(rng.uniform(-0.9, 0.9)), prec) if i % 3 else zero for i in range(n)])
    P = np.stack([qb.from_str(repr(rng.uniform(-0.9, 0.9)), prec) if i % 3 else zero for i in range(n)])
    return spin, delta, S, P


@pytest.mark.parametrize("prec,n_max,lam,dim,n", [(600, 100, 11, "3", 256), (800, 200, 13, "3", 256), (1024, 400, 19, "3", 256),
                                                  (1536, 800, 27, "3", 256), (1000, 150, 14, "4", 64), (600, 60, 8, "5/2", 64),
                                                  (1200, 100, 10, "3", 64)])
def test_tables_vs_reference_live(ref, prec, n_max, lam, dim, n):
    """BASELINE configs C1-C4 (and other dims / precisions) on seeded inputs, against the reference run on this box"""
    rng = random.Random(prec * 31 + lam)
    spin, delta, S, P = _random_inputs(rng, prec, n)
    eng = _engine(prec, n_max, lam, dim)
    got = eng.gblock(spin, delta, S, P)
    want = ref.gblock(prec, n_max, lam, dim, delta, spin, S, P)
    bad = [i for i in range(n) if not np.array_equal(got[i], want[i])]
    assert not bad, bad
    # F / H blocks over the resident tables
    d = np.stack([qb.from_str("0.5181489", prec), qb.from_str("0.96543695", prec), qb.from_str("1.412625", prec)])
    tix = [rng.randrange(n) for _ in range(12)]
    dix = [rng.randrange(3) for _ in range(12)]
    sym = [rng.randrange(2) for _ in range(12)]
    fb = eng.fblock(tix, d, dix, sym)
    for j in range(12):
        w = ref.fblock(prec, n_max, lam, dim, d[dix[j]], delta[tix[j]], int(spin[tix[j]]), S[tix[j]], P[tix[j]], sym[j])
        assert np.array_equal(fb[j], w), j
    eng.close()


def test_special_operators_vs_reference(ref):
    """unit operator, operators at the three families of divergent points, and their neighbours"""
    prec, n_max, lam, dim = 1024, 120, 9, "3"
    eng = _engine(prec, n_max, lam, dim)
    cases = [(0, "0"), (2, "3"), (0, "1/2"), (0, "1"), (1, "2"), (3, "4"), (4, "5"), (0, "0.25"), (2, "3.0000001"),
             (1, "1"), (5, "6"), (0, "3/2")]
    spin = np.array([c[0] for c in cases], dtype=np.uint32)
    delta = np.stack([qb.from_str(c[1], prec) for c in cases])
    Z = qb.zeros(len(cases), prec)
    got = eng.gblock(spin, delta, Z, Z)
    want = ref.gblock(prec, n_max, lam, dim, delta, spin, Z, Z)
    bad = [cases[i] for i in range(len(cases)) if not same_records(want[i], got[i])]
    assert not bad, bad
    eng.close()


def test_full_size_properties():
    """north-star size (lambda = 19, n_Max = 400, 1024-bit): size-independent properties of a PMP-sized batch"""
    prec, n_max, lam = 1024, 400, 19
    eng = _engine(prec, n_max, lam)
    rng = random.Random(5)
    n = 300
    spin, delta, S, P = _random_inputs(rng, prec, n, spins=tuple(range(0, 51)))
    delta[7] = qb.from_fraction(0, prec)  # unit operator
    spin[7] = 0
    S[7] = P[7] = 0
    g1 = eng.gblock(spin, delta, S, P)
    # (1) results do not depend on batch composition or position (the reference's thread-count independence)
    perm = np.array(rng.sample(range(n), n))
    g2 = eng.gblock(spin[perm], delta[perm], S[perm], P[perm])
    assert np.array_equal(g2, g1[perm])
    sub = list(range(0, n, 17))
    g3 = eng.gblock(spin[sub], delta[sub], S[sub], P[sub])
    assert np.array_equal(g3, g1[sub])
    # (2) idempotence
    assert np.array_equal(eng.gblock(spin, delta, S, P), g1)
    # (3) unit operator: g = (4 rho)^0 * b0 = 1 on the diagonal, derivatives vanish (src/hor_recursion.cpp:20)
    one = qb.from_fraction(1, prec)
    assert np.array_equal(g1[7, 0], one)
    assert all((int(w[0]) & 3) == 0 for w in g1[7, 1:])
    # (4) every record is canonical: regular numbers are normalised and rounded to prec bits
    top = g1[..., -1]
    kind = g1[..., 0] & 3
    assert np.all((kind == 0) | ((kind == 1) & ((top >> 31) == 1)))
    # (5) F = 2 proj_odd(v^d g) is linear in g only through exact doubling: F(d, g) computed twice is identical and the
    #     Even / Odd blocks together hold dim_M entries
    d = qb.from_str("0.5181489", prec).reshape(1, -1)
    fo = eng.fblock([3, 3], d, [0, 0], [1, 0])
    assert len(fo[0]) + len(fo[1]) == eng.dim_M
    eng.close()


def test_error_behaviour():
    eng = qb.Engine(1024)
    with pytest.raises(qb.QbError):
        eng.lib.qb_gblock_run(eng.h) and None
        eng._check(eng.lib.qb_gblock_run(eng.h), "run")  # no context yet
    assert eng.lib.qb_gblock_run(eng.h) == -3
    eng.context(50, 5)
    assert eng.gblock([], np.zeros((0, eng.W), np.uint32), np.zeros((0, eng.W), np.uint32), np.zeros((0, eng.W), np.uint32)).shape[0] == 0
    with pytest.raises(qb.QbError):
        qb.Engine(4096)  # no kernels for this precision
    eng.close()


@pytest.mark.parametrize("prec", [512, 600, 800, 1024, 1200, 1536])
def test_device_fast_paths_vs_mpfr(ref, prec):
    """the register-resident routines with their PTX carry chains (fastops.cuh), on the device, against MPFR"""
    from tests.test_host_arith import FAST, fast_cases

    eng = _engine(prec, 20, 5)
    for name, kw in fast_cases(ref, prec, 400, prec + 1):
        want = ref.op(name, prec, **kw)
        got = eng.op(FAST[name], **kw)
        bad = [i for i in range(len(want)) if not np.array_equal(want[i], got[i])]
        assert not bad, (name, bad[:5])
    eng.close()


def test_chunked_streams_match_single_chunk(tmp_path):
    """Large batches are cut into chunks, and chunks into front groups, spread over several streams (coefficients + series +
    powers per group on one stream, scale + Horner per chunk on a second, the transverse recursion on a third, finished
    chunks copied out on a copy stream).  Scheduling must be invisible in the output: a child process with a tiny chunk and
    group size (QB_CHUNK / QB_FRONT are read once per process) must produce the tables of the single-chunk run bit for
    bit, through both the one-shot call and plan/run/fetch."""
    import os
    import subprocess
    import sys

    prec, n_max, lam, n = 600, 100, 11, 1500
    rng = random.Random(11)
    spin, delta, S, P = _random_inputs(rng, prec, n)
    delta[5] = qb.from_fraction(Fraction(1, 2), prec)  # divergent for spin 0 in d = 3: the table owns two series jobs
    spin[5] = 0
    np.savez(tmp_path / "in.npz", spin=spin, delta=delta, S=S, P=P)
    code = (
        "import numpy as np, qboot_b200 as qb\n"
        f"z = np.load(r'{tmp_path / 'in.npz'}')\n"
        f"eng = qb.Engine({prec}).context({n_max}, {lam}, '3')\n"
        "g = eng.gblock(z['spin'], z['delta'], z['S'], z['P'])\n"
        "eng.plan(z['spin'], z['delta'], z['S'], z['P']); eng.run(); eng.sync(); g2 = eng.fetch()\n"
        f"np.savez(r'{tmp_path / 'out.npz'}', g=g, g2=g2, launches=eng.launches)\n"
    )
    env = dict(os.environ, QB_CHUNK="256", QB_FRONT="600", PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    subprocess.run([sys.executable, "-c", code], check=True, env=env, timeout=600)
    out = np.load(tmp_path / "out.npz")
    eng = _engine(prec, n_max, lam)
    want = eng.gblock(spin, delta, S, P)
    eng.close()
    assert int(out["launches"]) > 2 * 6 * 5  # several chunks really ran
    assert np.array_equal(out["g"], want)
    assert np.array_equal(out["g2"], want)


@pytest.mark.parametrize("prec,n_max,lam,n", [(600, 100, 11, 160), (800, 120, 9, 160), (1024, 150, 9, 320), (1200, 90, 8, 96), (1536, 100, 9, 96)])
def test_series_kernels_agree(ref, monkeypatch, prec, n_max, lam, n):
    """The series recurrence has two kernels: one thread per job (k_series) and, for the batches of a single PMP, eight warps
    per 32 jobs (k_series_coop: products and polynomials off the chain).  QB_SERIES_COOP (read at every launch) chooses; the
    same batch through both must give the same tables bit for bit, equal to the reference's -- special operators, zero S / P
    (products with a zero operand leave the frame hand-over) and both recurrence lengths (spin 0 / spin > 0) included."""
    rng = random.Random(77 + prec)
    spin, delta, S, P = _random_inputs(rng, prec, n)
    for i, (sp, d) in enumerate([(0, "0"), (0, "1/2"), (2, "3"), (1, "2"), (0, "1"), (3, "4")]):
        spin[i] = sp
        delta[i] = qb.from_str(d, prec)
    eng = _engine(prec, n_max, lam)
    monkeypatch.setenv("QB_SERIES_COOP", "0")
    per_thread = eng.gblock(spin, delta, S, P)
    monkeypatch.setenv("QB_SERIES_COOP", "64")
    coop = eng.gblock(spin, delta, S, P)
    eng.close()
    assert np.array_equal(per_thread, coop)
    want = ref.gblock(prec, n_max, lam, "3", delta, spin, S, P)
    bad = [i for i in range(n) if not same_records(want[i], coop[i])]
    assert not bad, bad


def test_sink_receives_the_tables_of_every_run():
    """qb_gblock_set_sink: every finished chunk of a run is copied to the sink (here page-locked host memory; in a multi-GPU
    job a peer GPU's buffer opened with qb_ipc_open) and the handle's stream joins the copies"""
    import ctypes as C

    prec, n_max, lam, n = 600, 60, 9, 700
    rng = random.Random(5)
    spin, delta, S, P = _random_inputs(rng, prec, n)
    eng = _engine(prec, n_max, lam)
    nbytes = n * eng.dim_M * eng.W * 4
    buf = eng.lib.qb_pinned_alloc(nbytes)
    assert buf
    try:
        view = np.frombuffer((C.c_ubyte * nbytes).from_address(buf), dtype=np.uint32).reshape(n, eng.dim_M, eng.W)
        view[:] = 0
        eng._check(eng.lib.qb_gblock_set_sink(eng.h, buf), "qb_gblock_set_sink")
        eng.plan(spin, delta, S, P)
        eng.run()
        eng.sync()
        want = eng.fetch()
        assert np.array_equal(view, want)
        # switched off: a later run leaves the buffer alone
        eng._check(eng.lib.qb_gblock_set_sink(eng.h, None), "qb_gblock_set_sink")
        view[:] = 0
        eng.run()
        eng.sync()
        assert not view.any()
    finally:
        eng.lib.qb_gblock_set_sink(eng.h, None)
        eng.close()
        eng.lib.qb_pinned_free(buf)


def test_poly_eval_batch_vs_reference(ref):
    """qb_poly_eval_batch -- the bilinear bases at the sample points: Horner by fused multiply-adds (Polynomial::eval with a
    real argument, polynomial.hpp:66-83) times a scale -- against mpfr_fma / mpfr_mul chains"""
    import ctypes as C

    from tests.golden.make_golden import rnd_num
    from tests.util import P

    prec = 1000
    eng = _engine(prec, 20, 5)
    rng = random.Random(77)
    degs = [-1, 0, 1, 2, 5, 13, 24]                      # -1: the zero polynomial
    coefs = [[rnd_num(rng, prec, -8, 8) for _ in range(d + 1)] for d in degs]
    ofs = np.cumsum([0] + [d + 1 for d in degs]).astype(np.int32)
    flat = np.stack([c for cs in coefs for c in cs])
    xs = np.stack([rnd_num(rng, prec, -2, 6) for _ in range(9)])
    ss = np.stack([rnd_num(rng, prec, -3, 3) for _ in range(4)])
    items = np.array([(p, x, s) for p in range(len(degs)) for x in range(len(xs)) for s in range(len(ss))], dtype=np.int32)
    out = np.zeros((len(items), eng.W), dtype=np.uint32)
    eng._check(eng.lib.qb_poly_eval_batch(eng.h, len(degs), P(ofs), P(flat), len(xs), P(xs), len(ss), P(ss), len(items), P(items), P(out)),
               "qb_poly_eval_batch")
    zero = qb.from_fraction(0, prec)
    for i, (p, x, s) in enumerate(items):
        if degs[p] < 0:
            acc = zero
        else:
            acc = coefs[p][degs[p]]
            for k in range(degs[p] - 1, -1, -1):
                acc = ref.op("fma", prec, a=acc, b=xs[x], c=coefs[p][k])[0]
        want = ref.op("mul", prec, a=acc, b=ss[s])[0]
        assert same_records(want[None, :], out[i][None, :]), (p, x, s)
    eng.close()


def test_bench_workload_vs_reference(ref):
    """The benchmark's own workload (mixed-Ising key set at lambda = 19, n_Max = 400, 1024-bit, full-width sample points)
    sampled across spins and (S, P) classes, against the reference run on this box: what bench.py times is what is pinned."""
    from qboot_b200 import workload as wl

    prec, n_max, lam = 1024, 400, 19
    keys = wl.sweep_keys(prec, lam, 20, 50, 49096)[::89]  # 552 tables across every sector, spin and (S, P) class
    spin, delta, S, P = wl.to_records(keys, prec)
    eng = _engine(prec, n_max, lam)
    got = eng.gblock(spin, delta, S, P)
    eng.close()
    want = ref.gblock(prec, n_max, lam, "3", delta, spin, S, P)
    bad = [i for i in range(len(keys)) if not np.array_equal(got[i], want[i])]
    assert len(keys) >= 512 and not bad, bad


def test_pow_second_ziv_level_is_correctly_rounded(ref):
    """x^y is exp(y ln x) at 96 guard bits with a second level at 192 when the rounding is in doubt.  A real doubt has
    probability 2^-40 per power, so the retry is forced by construction: the in-doubt window is widened until about one
    power in eight takes the second level; results must still be MPFR's (mpfr_pow is correctly rounded), the counters must
    show that the second level ran and that nothing stayed in doubt."""
    from tests.golden.make_golden import rnd_num

    for prec, n_max, lam in ((1024, 60, 9), (600, 40, 7)):
        eng = _engine(prec, n_max, lam)
        guard = 32 * (qb.nlimbs(prec) + 3) - prec
        rng = random.Random(prec)
        ys = np.stack([qb.from_fraction(Fraction(rng.randint(1, 10**8), rng.randint(10**5, 10**6)), prec) for _ in range(400)])
        rho = np.tile(eng.context_consts()[0], (len(ys), 1))
        want_rho = ref.op("pow", prec, a=rho, b=ys)
        want_4 = ref.op("ui_pow", prec, a=ys, ia=np.full(len(ys), 4, dtype=np.int64))
        assert eng.ziv_counters() == (0, 0)
        assert np.array_equal(eng.op(OPS["pow"], b=ys), want_rho) and np.array_equal(eng.op(OPS["ui_pow"], a=ys), want_4)
        assert eng.ziv_counters() == (0, 0)          # default window: no power of this sample is in doubt
        eng.set_ziv_slack(guard - 4)                 # in doubt when the three bits below the round bit all differ from it
        assert np.array_equal(eng.op(OPS["pow"], b=ys), want_rho) and np.array_equal(eng.op(OPS["ui_pow"], a=ys), want_4)
        redone, undecided = eng.ziv_counters()
        assert 40 <= redone <= 200 and undecided == 0, (redone, undecided)
        # the table pipeline (k_pow: 4^Delta and rho^(Delta - k), k = 0..lambda) under the same window
        spin = np.array([0, 1, 2, 3, 4, 6], dtype=np.uint32)
        delta = np.stack([qb.from_fraction(Fraction(rng.randint(10**6, 10**8), 10**6), prec) for _ in spin])
        z = qb.zeros(len(spin), prec)
        got = eng.gblock(spin, delta, z, z)
        assert eng.ziv_counters()[0] > redone and eng.ziv_counters()[1] == 0
        assert np.array_equal(got, ref.gblock(prec, n_max, lam, "3", delta, spin, z, z))
        eng.close()


# ---- the PMP half of the path: planner + device assembly / pack / decimal text through the C++ facade -------------------
def _same_dirs(a, b):
    import filecmp
    import os

    fa, fb = sorted(os.listdir(a)), sorted(os.listdir(b))
    assert fa == fb
    return [f for f in fa if not filecmp.cmp(os.path.join(a, f), os.path.join(b, f), shallow=False)]


@pytest.mark.parametrize("prec,n_max,lam,model,mode,finite,numax,maxspin", [
    (600, 100, 11, 0, 0, True, 8, 24),      # BASELINE config 1: single-correlator Ising, 600-bit, spins 0-24 + finite range
    (800, 200, 13, 1, 1, True, 10, 30),     # BASELINE config 2: mixed Ising, Delta12 != 0 blocks, finite ranges [lb, ub), 800-bit
    (1000, 120, 10, 1, 0, False, 6, 12),    # FindContradiction on the mixed system, 1000 bits (not a limb multiple)
])
def test_sdpb_directory_equals_reference(ref, tmp_path, prec, n_max, lam, model, mode, finite, numax, maxspin):
    """convert -> create_input -> write on the GPU engine vs the reference's MPFR build: every file byte for byte"""
    want, got = str(tmp_path / "ref"), str(tmp_path / "ours")
    r = ref.build_pmp(prec, n_max, lam, "3", model, mode, finite, numax=numax, maxspin=maxspin, threads=8, outdir=want)
    o = qb.build_pmp(prec, n_max, lam, "3", model, mode, finite, numax=numax, maxspin=maxspin, threads=8, outdir=got)
    assert o["constraints"] == r["constraints"] and o["N"] == r["N"]
    assert _same_dirs(got, want) == []


def test_device_decimal_text_vs_reference(ref):
    """k_dec_format against the reference's ostream (mpfr_get_str digits + default-float layout) at the SDPB digit count"""
    import ctypes as C

    from tests.golden.make_golden import rnd_num
    from tests.util import P

    for prec in (600, 1024, 1536):
        eng = _engine(prec, 20, 5)
        rng = random.Random(prec)
        nd = 3 + int(prec * 0.302)
        vals = [rnd_num(rng, prec, -500, 500) for _ in range(300)]  # |x| < 10^151 < 10^nd: inside the device formatter's range
        vals += [ref.from_str(s, prec) for s in ["1", "0.5", "10", "1000", "999999", "1e-5", "0.001", "0.0001", "-2", "123.456", "0.1", "0"]]
        vals.append(ref.op("add", prec, a=ref.from_str("0.5", prec), b=ref.from_int_exp(0, 1, -(nd + 1), prec))[0])  # a tie
        vals += [ref.from_str("1e1500", prec), ref.from_str("-3e-2000", prec)]  # outside the device range: declined
        x = np.stack(vals)
        slot = nd + 16
        slots = np.zeros((len(x), slot), dtype=np.uint8)
        ln = np.zeros(len(x), dtype=np.int32)
        eng._check(eng.lib.qb_dec_format(eng.h, nd, len(x), P(x), P(slots), slot, P(ln)), "qb_dec_format")
        declined = []
        for i, v in enumerate(x):
            if ln[i] < 0:
                declined.append((i, int(ln[i]), ref.to_str(v, prec, 12)))
                continue
            assert slots[i, : ln[i]].tobytes().decode() == ref.to_str(v, prec, nd) + "\n", ref.to_str(v, prec, 20)
        assert [d[0] for d in declined] == [len(x) - 2, len(x) - 1], declined
        eng.close()


def _run_driver(exe, cwd, *args):
    import os
    import subprocess

    os.makedirs(cwd, exist_ok=True)
    subprocess.run([exe, *args], cwd=cwd, check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900)
    (d,) = [e for e in os.listdir(cwd) if os.path.isdir(os.path.join(cwd, e))]
    return d, os.path.join(cwd, d)


@pytest.mark.parametrize("ours,theirs,args", [
    ("sample_b200", "sample_ref", ()),             # sample/main.cpp: single-correlator Ising, 1000-bit, n_Max 400, lambda 16
    ("testmain_b200", "testmain_ref", ("solve",)),  # test/src/main.cpp solve: mixed Ising, ExtremalOPE, lambda 14
])
def test_reference_drivers_run_unmodified_on_the_facade(tmp_path, ours, theirs, args):
    """The drop-in boundary, demonstrated: the reference's own driver programs compiled UNMODIFIED against include/qboot and
    linked with qboot_b200/libqboot.so (CUDA engine) write the same SDPB directory, byte for byte, as the same sources
    linked with the reference (built by __graft_entry__.build() / oracle/Makefile where /root/reference exists)."""
    import os

    from tests.conftest import ROOT

    a, b = os.path.join(ROOT, "tests", "_build", ours), os.path.join(ROOT, "oracle", "_ref", theirs)
    if not (os.path.exists(a) and os.path.exists(b)):
        pytest.skip("driver binaries not built (they need /root/reference at build time)")
    na, da = _run_driver(a, str(tmp_path / "ours"), *args)
    nb, db = _run_driver(b, str(tmp_path / "ref"), *args)
    assert na == nb
    assert _same_dirs(da, db) == []


def test_facade_single_calls_vs_reference(ref):
    """Context::gBlock / evaluate through the facade's one-table path equal the batched engine results (smoke of the API)"""
    prec, n_max, lam = 600, 60, 9
    eng = _engine(prec, n_max, lam)
    d = qb.from_str("1.7", prec)
    z = qb.from_fraction(0, prec)
    t = eng.gblock([2], d, z, z)
    assert np.array_equal(t, ref.gblock(prec, n_max, lam, "3", d, [2], z, z))
    dv = np.stack([qb.from_str("0.5181489", prec)])
    out = np.zeros((1, eng.dim_M, eng.W), dtype=np.uint32)
    from tests.util import P
    eng._check(eng.lib.qb_vtod_batch(eng.h, 1, P(dv), P(out)), "qb_vtod_batch")
    fb = np.zeros((qb.load_library().qb_block_dim(eng.h, 1), eng.W), dtype=np.uint32)
    sym = np.array([1], dtype=np.int32)
    eng._check(eng.lib.qb_fblock_tables(eng.h, 1, P(np.ascontiguousarray(t[0])), P(dv), P(sym), P(fb)), "qb_fblock_tables")
    assert np.array_equal(fb, ref.fblock(prec, n_max, lam, "3", dv[0], d, 2, z, z, 1))
    eng.close()
