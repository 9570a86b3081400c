#!/usr/bin/env python
"""Probe (development aid): BASELINE config 4 (lambda = 27, n_Max = 800, 1536-bit) over several chunks, throughput and a
sample of tables against the reference."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import qboot_b200 as qb
from qboot_b200 import workload as wl
from oracle import ref

prec, n_max, lam, n = 1536, 800, 27, int(sys.argv[1]) if len(sys.argv) > 1 else 6000
keys = wl.sweep_keys(prec, lam, 20, 80, n)
spin, delta, S, P = wl.to_records(keys, prec)
eng = qb.Engine(prec).context(n_max, lam, "3")
t0 = time.time(); g = eng.gblock(spin, delta, S, P); t1 = time.time()
t0 = time.time(); g = eng.gblock(spin, delta, S, P); t1 = time.time()
print(f"C4: {n} tables in {t1 - t0:.3f} s end to end -> {n / (t1 - t0):.0f} tables/s; launches {eng.launches}", flush=True)
idx = np.arange(0, n, max(1, n // 12))[:12]
want = ref.gblock(prec, n_max, lam, "3", delta[idx], spin[idx], S[idx], P[idx])
print("sample of", len(idx), "tables bit-identical to the reference:", bool(np.array_equal(g[idx], want)))
