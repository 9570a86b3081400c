#!/usr/bin/env python
"""Smallest possible GPU invocation (for compute-sanitizer)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qboot_b200 as qb
prec, n_max, lam = 600, 12, 3
eng = qb.Engine(prec).context(n_max, lam, "3")
delta = np.stack([qb.from_str("1.7", prec), qb.from_str("4.25", prec)])
Z = qb.zeros(2, prec)
g = eng.gblock([0, 2], delta, Z, Z)
print("ok", g.shape, g[0, 0, :4])
