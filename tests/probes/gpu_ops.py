#!/usr/bin/env python
"""GPU primitive check (development aid): device build of each mpf routine vs MPFR through oracle/_ref."""
import os, sys, random
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import qboot_b200 as qb
from oracle import ref

rng = random.Random(1)
def rnd_num(prec, emin=-40, emax=40):
    style = rng.choice(['full', 'short', 'pow2', 'ones', 'zero'])
    if style == 'zero': return ref.from_int_exp(rng.randint(0, 1), 0, 0, prec)
    if style == 'full': m = rng.getrandbits(prec) | (1 << (prec - 1))
    elif style == 'short': m = rng.getrandbits(rng.randint(1, min(prec, 70))) | 1
    elif style == 'pow2': m = 1
    else:
        m = (1 << prec) - 1
        if rng.random() < 0.5: m ^= rng.getrandbits(3)
    e = rng.randint(emin, emax)
    return ref.from_int_exp(rng.randint(0, 1), m, e - m.bit_length(), prec)

for prec in [int(x) for x in (sys.argv[1:] or ["600", "1024"])]:
    eng = qb.Engine(prec).context(20, 5, "3")
    n = 300
    a = np.stack([rnd_num(prec) for _ in range(n)]); b = np.stack([rnd_num(prec) for _ in range(n)]); c = np.stack([rnd_num(prec) for _ in range(n)])
    b2 = b.copy()
    for i in range(0, n, 3):
        b2[i] = a[i]; b2[i, 2] ^= (rng.getrandbits(8) << ((-prec) % 32)) & 0xFFFFFFFF; b2[i, 0] ^= 0x80000000 if (a[i, 0] & 3) else 0
    ia = np.array([rng.choice([0, 1, -1, 2, -3, 7, 400, -12345, 2**31 - 1, 2**40 + 3, rng.randint(-10**6, 10**6)]) for _ in range(n)], dtype=np.int64)
    ib = np.array([rng.choice([1, 2, 3, 7, 10, 20, 1000, 2**32 - 1, rng.randint(1, 10**6)]) for _ in range(n)], dtype=np.int64)
    ia32 = np.clip(ia, -(2**32 - 1), 2**32 - 1)
    tests = [('add', dict(a=a, b=b)), ('add', dict(a=a, b=b2)), ('sub', dict(a=a, b=b2)), ('mul', dict(a=a, b=b)), ('div', dict(a=a, b=b)),
             ('div', dict(a=a, b=b2)), ('fma', dict(a=a, b=b, c=c)), ('mul_si', dict(a=a, ia=ia)), ('add_si', dict(a=a, ia=ia)),
             ('div_si', dict(a=a, ia=ia32)), ('si_sub', dict(a=a, ia=ia)), ('mul_q', dict(a=a, ia=ia, ib=ib)), ('div_q', dict(a=a, ia=ia32, ib=ib)),
             ('add_q', dict(a=a, ia=ia, ib=ib)), ('sub_q', dict(a=a, ia=ia, ib=ib)), ('set_q', dict(ia=ia, ib=ib)), ('cmp', dict(a=a, b=b2)), ('isint', dict(a=a))]
    for name, kw in tests:
        try:
            e = eng.op(ref.OPS[name], **kw)
        except Exception as ex:
            print(prec, name, 'EXCEPTION', ex, flush=True); sys.exit(1)
        r = ref.op(name, prec, **kw)
        if name in ('cmp', 'isint'): r = r[:, :1]; e = e[:, :1]
        bad = [i for i in range(len(r)) if not np.array_equal(r[i], e[i]) and not ((r[i, 0] & 3) >= 2 and (r[i, 0] & 3) == (e[i, 0] & 3))]
        print(prec, name, 'mismatches', len(bad), bad[:5], flush=True)
    eng.close()
