#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference (oracle/_ref).

Run in the build container (needs /root/reference compiled by `make -C oracle ref`):  python tests/golden/make_golden.py
The reference ships no golden vectors of its own, so these are outputs of the reference itself on
fixed seeded inputs; they pin the Python oracle, the host build of the engine's arithmetic and the CUDA path.
"""
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref  # noqa: E402


def rnd_num(rng, prec, emin=-40, emax=40):
    style = rng.choice(["full", "short", "pow2", "ones", "zero"])
    if style == "zero":
        return ref.from_int_exp(rng.randint(0, 1), 0, 0, prec)
    if style == "full":
        m = rng.getrandbits(prec) | (1 << (prec - 1))
    elif style == "short":
        m = rng.getrandbits(rng.randint(1, min(prec, 70))) | 1
    elif style == "pow2":
        m = 1
    else:
        m = (1 << prec) - 1
        if rng.random() < 0.5:
            m ^= rng.getrandbits(3)
    e = rng.randint(emin, emax)
    return ref.from_int_exp(rng.randint(0, 1), m, e - m.bit_length(), prec)


def golden_ops(prec, n=48, seed=2026):
    rng = random.Random(seed + prec)
    a = np.stack([rnd_num(rng, prec) for _ in range(n)])
    b = np.stack([rnd_num(rng, prec) for _ in range(n)])
    c = np.stack([rnd_num(rng, prec) for _ in range(n)])
    for i in range(0, n, 3):  # near-cancellation pairs
        b[i] = a[i]
        b[i, 2] ^= (rng.getrandbits(8) << ((-prec) % 32)) & 0xFFFFFFFF
        if a[i, 0] & 3:
            b[i, 0] ^= 0x80000000
    ia = np.array([rng.choice([0, 1, -1, 2, -3, 7, 400, -12345, 2**31 - 1, 2**40 + 3, rng.randint(-10**6, 10**6)])
                   for _ in range(n)], dtype=np.int64)
    ib = np.array([rng.choice([1, 2, 3, 7, 10, 20, 1000, 2**32 - 1, rng.randint(1, 10**6)]) for _ in range(n)],
                  dtype=np.int64)
    ia32 = np.clip(ia, -(2**32 - 1), 2**32 - 1)
    out = dict(a=a, b=b, c=c, ia=ia, ib=ib, ia32=ia32)
    for name, kw in [("add", dict(a=a, b=b)), ("sub", dict(a=a, b=b)), ("mul", dict(a=a, b=b)), ("div", dict(a=a, b=b)),
                     ("fma", dict(a=a, b=b, c=c)), ("mul_si", dict(a=a, ia=ia)), ("add_si", dict(a=a, ia=ia)),
                     ("div_si", dict(a=a, ia=ia32)), ("si_sub", dict(a=a, ia=ia)), ("mul_q", dict(a=a, ia=ia, ib=ib)),
                     ("div_q", dict(a=a, ia=ia32, ib=ib)), ("add_q", dict(a=a, ia=ia, ib=ib)),
                     ("sub_q", dict(a=a, ia=ia, ib=ib)), ("set_q", dict(ia=ia, ib=ib))]:
        out["r_" + name] = ref.op(name, prec, **kw)
    pos = a.copy()
    pos[:, 0] &= 0x7FFFFFFF
    out["r_sqrt"] = ref.op("sqrt", prec, a=pos)
    # pow(rho, y), 4^y for moderate y
    ys = np.stack([ref.from_str(repr(rng.uniform(-30, 300)), prec) for _ in range(n)])
    rho = ref.op("si_sub", prec, a=ref.op("sqrt", prec, a=ref.from_int_exp(0, 8, 0, prec)), ia=[3])[0]
    out["ys"] = ys
    out["r_pow_rho"] = ref.op("pow", prec, a=np.tile(rho, (n, 1)), b=ys)
    out["r_pow_4"] = ref.op("ui_pow", prec, a=ys, ia=[4] * n)
    np.savez_compressed(os.path.join(HERE, f"ops_p{prec}.npz"), **out)


def golden_tables(name, prec, n_max, lam, dim, n, seed, with_series=True):
    rng = random.Random(seed)
    spins = [rng.choice([0, 1, 2, 3, 8, 25]) for _ in range(n)]
    delta = np.stack([ref.from_str(repr(rng.uniform(0.6, 40.0)), prec) for _ in range(n)])
    zero = ref.from_int_exp(0, 0, 0, prec)
    S = np.stack([ref.from_str(repr(rng.uniform(-1, 1)), prec) if i % 2 else zero for i in range(n)])
    P = np.stack([ref.from_str(repr(rng.uniform(-1, 1)), prec) if i % 2 else zero for i in range(n)])
    # unit operator; stress tensor at the unitarity bound (divergent in d = 3)
    delta[0], spins[0], S[0], P[0] = zero, 0, zero, zero
    if dim == "3" and n > 2:
        delta[1], spins[1], S[1], P[1] = ref.from_int_exp(0, 3, 0, prec), 2, zero, zero
    spins = np.array(spins, dtype=np.uint32)
    g = ref.gblock(prec, n_max, lam, dim, delta, spins, S, P)
    out = dict(prec=prec, n_max=n_max, lam=lam, dim=dim, spin=spins, delta=delta, S=S, P=P, g=g)
    rho, mat = ref.context_consts(prec, n_max, lam, dim)
    out["rho"], out["mat"] = rho, mat
    if with_series:
        out["hblock"] = np.stack([ref.hblock(prec, n_max, dim, delta[i], int(spins[i]), S[i], P[i]) for i in range(n)])
        out["rec"] = [ref.rec_coeffs(prec, dim, delta[i], int(spins[i]), S[i], P[i]) for i in range(n)]
        out["rec"] = np.stack([np.pad(r, ((0, 8 - r.shape[0]), (0, 5 - r.shape[1]), (0, 0))) for r in out["rec"]])
    d = np.stack([ref.from_str("0.5181489", prec), ref.from_str("0.96543695", prec)])
    out["fb_d"] = d
    for sym in (0, 1):
        out[f"fb_sym{sym}"] = np.stack(
            [ref.fblock(prec, n_max, lam, dim, d[sym], delta[i], int(spins[i]), S[i], P[i], sym) for i in range(n)])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


if __name__ == "__main__":
    golden_ops(600)
    golden_ops(1024)
    golden_tables("tables_c1", 600, 100, 11, "3", 6, seed=11)
    golden_tables("tables_small", 256, 24, 5, "3", 5, seed=12)
    golden_tables("tables_fracdim", 600, 40, 6, "37/10", 4, seed=13)
    golden_tables("tables_c3", 1024, 400, 19, "3", 3, seed=14, with_series=False)
    print("golden vectors written")
