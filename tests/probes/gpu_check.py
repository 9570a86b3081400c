#!/usr/bin/env python
"""Quick GPU parity + timing probe (development aid): engine vs oracle/_ref on a few tables per configuration."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import qboot_b200 as qb  # noqa: E402
from oracle import ref  # noqa: E402


def main():
    import random

    rng = random.Random(11)
    cfgs = [(600, 100, 11, "3", 8), (800, 200, 13, "3", 8), (1024, 400, 19, "3", 8), (1536, 800, 27, "3", 4),
            (600, 100, 10, "37/10", 6)]
    if len(sys.argv) > 1:
        cfgs = cfgs[: int(sys.argv[1])]
    for prec, n_max, lam, dim, n in cfgs:
        eng = qb.Engine(prec).context(n_max, lam, dim)
        rho, mat = eng.context_consts()
        rrho, rmat = ref.context_consts(prec, n_max, lam, dim)
        print(f"[{prec},{n_max},{lam},{dim}] consts rho={np.array_equal(rho, rrho)} mat={np.array_equal(mat, rmat)}", flush=True)
        spins = [rng.choice([0, 1, 2, 3, 8, 25]) for _ in range(n)]
        delta = np.stack([qb.from_str(repr(rng.uniform(0.6, 40.0)), prec) for _ in range(n)])
        S = np.stack([qb.from_str(repr(rng.uniform(-1, 1)), prec) if i % 2 else qb.from_fraction(0, prec) for i in range(n)])
        P = np.stack([qb.from_str(repr(rng.uniform(-1, 1)), prec) if i % 2 else qb.from_fraction(0, prec) for i in range(n)])
        # special operators: unit, stress tensor at the unitarity bound (divergent in d=3: Delta = 3, spin 2)
        delta[0] = qb.from_fraction(0, prec); spins[0] = 0; S[0] = 0; P[0] = 0
        if dim == "3":
            delta[1] = qb.from_fraction(3, prec); spins[1] = 2; S[1] = 0; P[1] = 0
        t0 = time.time(); g = eng.gblock(spins, delta, S, P); t1 = time.time()
        r = ref.gblock(prec, n_max, lam, dim, delta, spins, S, P); t2 = time.time()
        bad = [i for i in range(n) if not np.array_equal(g[i], r[i])]
        print(f"   gblock n={n}: gpu {t1 - t0:.3f}s ref {t2 - t1:.3f}s mismatching tables: {bad}", flush=True)
        d = np.stack([qb.from_str("0.5181489", prec), qb.from_str("0.96543695", prec)])
        tix = list(range(n)) * 2
        dix = [0] * n + [1] * n
        sym = [1] * n + [0] * n
        fb = eng.fblock(tix, d, dix, sym)
        badf = []
        for j, (ti, di, sy) in enumerate(zip(tix, dix, sym)):
            fr = ref.fblock(prec, n_max, lam, dim, d[di], delta[ti], spins[ti], S[ti], P[ti], sy)
            if not np.array_equal(fr, fb[j]):
                badf.append(j)
        print(f"   fblock terms={len(tix)} mismatching: {badf}", flush=True)
        eng.close()


if __name__ == "__main__":
    main()
