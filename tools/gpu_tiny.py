"""Small end-to-end exercise of every kernel (development aid; the target of `compute-sanitizer --tool memcheck`):
a few hundred tables incl. a unit and a divergent operator, F/H blocks, and one small PMP build through the facade."""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qboot_b200 as qb  # noqa: E402
from qboot_b200 import workload as wl  # noqa: E402

prec, n_max, lam = 600, 40, 7
eng = qb.Engine(prec).context(n_max, lam, "3")
keys = wl.mixed_ising_keys(prec, lam, 4, 8)[:300]
spin, delta, S, P = wl.to_records(keys, prec)
delta[3] = qb.from_fraction(0, prec)                      # unit operator
delta[5] = qb.from_str("0.5", prec)                       # divergent for spin 0 in d = 3
spin[3] = spin[5] = 0
g = eng.gblock(spin, delta, S, P)
fb = eng.fblock(list(range(16)), np.stack([qb.from_str("0.5181489", prec)]), [0] * 16, [i % 2 for i in range(16)])
print("tables", g.shape, "fblocks", len(fb), "launches", eng.launches)
eng.close()
d = tempfile.mkdtemp(prefix="qb_tiny_")
r = qb.build_pmp(prec, n_max, lam, "3", 1, 1, True, numax=3, maxspin=4, threads=2, outdir=d)
print("pmp", r["constraints"], r["N"], len(os.listdir(d)), "files")
