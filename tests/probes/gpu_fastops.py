#!/usr/bin/env python
"""GPU check of the register-resident fast paths (PTX carry chains) against MPFR through oracle/_ref."""
import os, sys, random
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import qboot_b200 as qb
from oracle import ref
from tests.golden.make_golden import rnd_num

rng = random.Random(3)
for prec in [int(x) for x in (sys.argv[1:] or ["1024"])]:
    eng = qb.Engine(prec).context(20, 5, "3")
    n = 600
    a = np.stack([rnd_num(rng, prec, -5, 5) for _ in range(n)]); b = np.stack([rnd_num(rng, prec, -5, 5) for _ in range(n)]); c = np.stack([rnd_num(rng, prec, -80, 80) for _ in range(n)])
    prod = ref.op('mul', prec, a=a, b=b)
    L = (prec + 31) // 32
    for i in range(0, n, 2):
        c[i] = prod[i].copy()
        if (c[i, 0] & 3) == 1:
            c[i, 0] ^= 0x80000000
            if rng.random() < 0.8:
                li = rng.randrange(L); bit = rng.randrange(31 if li == L - 1 else 32); c[i, 2 + li] ^= (1 << bit)
            c[i, 2] &= (0xFFFFFFFF << ((-prec) % 32)) & 0xFFFFFFFF
            c[i, 1] = np.uint32((int(np.int32(c[i, 1])) + rng.choice([0, 0, 1, -1, 31, 32, 33, 63, 64, 65, -31, -32, -33, -64, -65, prec - 33, prec, prec + 1, prec + 33, prec + 64, prec + 65, 3000, -prec, -prec - 1, -prec - 65, -3000, rng.randint(-prec - 80, prec + 80)])) & 0xFFFFFFFF)
    ia = np.array([rng.choice([1, -1, 2, 3, 7, 400, -12345, 2**31 - 1, 2**32 - 1, -(2**32 - 1)]) for _ in range(n)], dtype=np.int64)
    b2 = b.copy()
    for i in range(0, n, 3):
        b2[i] = a[i]; li = rng.randrange(L); b2[i, 2 + li] ^= (1 << rng.randrange(31 if li == L - 1 else 32)); b2[i, 2] &= (0xFFFFFFFF << ((-prec) % 32)) & 0xFFFFFFFF
    ok = [i for i in range(n) if (b2[i, 0] & 3) == 1 and (b[i, 0] & 3) == 1]
    for name, op, kw in [('mul', 40, dict(a=a, b=b)), ('fma', 41, dict(a=a, b=b, c=c)), ('add', 42, dict(a=prod, b=c)), ('sub', 43, dict(a=prod, b=c)),
                         ('add', 42, dict(a=a, b=b)), ('mul_si', 44, dict(a=a, ia=ia)), ('add_si', 45, dict(a=a, ia=ia)), ('add_si', 45, dict(a=c, ia=ia)),
                         ('div', 46, dict(a=a[ok], b=b[ok])), ('div', 46, dict(a=a[ok], b=b2[ok]))]:
        r = ref.op(name, prec, **kw); e = eng.op(op, **kw)
        m = [i for i in range(len(r)) if not np.array_equal(r[i], e[i])]
        print(prec, name, op, 'mismatches', len(m), m[:6], flush=True)
        if m:
            i = m[0]
            for k, v in kw.items(): print('   ', k, [hex(int(x)) for x in (v[i][:3] if v.ndim > 1 else [v[i]])], [hex(int(x)) for x in v[i][-2:]] if v.ndim > 1 else '')
            print('    ref', [hex(int(x)) for x in r[i][:4]], [hex(int(x)) for x in r[i][-2:]]); print('    got', [hex(int(x)) for x in e[i][:4]], [hex(int(x)) for x in e[i][-2:]])
    eng.close()
