#!/usr/bin/env python
"""GPU timing probe (development aid): per-kernel device times vs batch size at the north-star size."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qboot_b200 as qb
from qboot_b200 import workload as wl

prec, n_max, lam = 1024, 400, 19
sizes = [int(x) for x in (sys.argv[1:] or ["1292", "6137", "20000"])]
eng = qb.Engine(prec).context(n_max, lam, "3")
peak = eng.imad_peak(4096)
print(f"imad peak {peak:.3e} limb-MAC/s", flush=True)
eng.set_timing(True)
keys = wl.sweep_keys(prec, lam, 20, 50, max(sizes))
t0 = time.time(); spin, delta, S, P = wl.to_records(keys, prec); print(f"records for {len(keys)} keys in {time.time()-t0:.1f}s", flush=True)
for n in sizes:
    t0 = time.time(); eng.plan(spin[:n], delta[:n], S[:n], P[:n]); t1 = time.time()
    for rep in range(2):
        eng.run(); ms = eng.kernel_times()
    tot = sum(ms)
    w = sum(wl.work_per_table(n_max, lam, 32, int(s))["total"] for s in spin[:n])
    wd = sum(wl.work_per_table(n_max, lam, 32, int(s))["deriv"] for s in spin[:n])
    print(f"n={n}: plan {t1-t0:.2f}s kernels ms rec={ms[0]:.1f} series={ms[1]:.1f} deriv={ms[2]:.1f} finish={ms[3]:.1f} total={tot:.1f} "
          f"-> {n/tot*1e3:.0f} tables/s, {w/tot*1e3/peak*100:.1f}% of IMAD peak (deriv kernel {wd/ms[2]*1e3/peak*100:.1f}%)", flush=True)
    eng.set_timing(False); eng.run(); eng.sync(); t0 = time.time()
    for rep in range(3): eng.run()
    eng.sync(); dt = (time.time() - t0) / 3; eng.set_timing(True)
    print(f"   untimed (chunked, QB_CHUNK={os.environ.get('QB_CHUNK','default')}): {dt*1e3:.1f} ms -> {n/dt:.0f} tables/s, {w/dt/peak*100:.1f}% of IMAD peak", flush=True)
    t0 = time.time(); out = eng.fetch(); print(f"   fetch {time.time()-t0:.2f}s {out.nbytes/1e6:.0f} MB", flush=True)
