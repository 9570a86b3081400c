"""Synthetic workloads shaped like the reference's PMP builds (deterministic, no RNG).

The key set of a PMP build is defined by the reference, not random: for a continuous sector
with spin l the block tables are needed at the sample points Delta_k = x_k + gap, k = 0..D, with
D = num_poles(l) + lambda, num_poles(l) = numax + min(numax, l) / 2 (include/qboot/bootstrap_equation.hpp:32-40),
x_k = pi^2 (4k - 1)^2 / (-64 (D + 1) ln rho*) (include/qboot/conformal_scale.hpp:105-116) and gap = the unitarity
bound (include/qboot/primary_op.hpp:18-21) or the model's gap.  The x_k are evaluated here in double precision and
converted exactly: as inputs they are as good as the reference's (any Delta is a valid operator dimension); parity
tests feed identical records to both sides.
"""
from __future__ import annotations

import functools
import math
from fractions import Fraction

import numpy as np

from . import from_fraction

LN_RHO = math.log(3.0 - 2.0 * math.sqrt(2.0))


def num_poles(numax: int, spin: int) -> int:
    return numax + min(numax, spin) // 2


@functools.lru_cache(maxsize=None)
def _sample_points(D: int):
    """x_k = pi^2 (4k-1)^2 / (-64 (D+1) ln(3 - 2 sqrt 2)), k = 0..D (conformal_scale.hpp:105-116) to 1600 bits, as exact
    rationals: like the reference's, the sample points carry full-width mantissas at every supported precision"""
    import mpmath

    with mpmath.workprec(1600):
        c = mpmath.pi ** 2 / (-64 * (D + 1) * mpmath.log(3 - 2 * mpmath.sqrt(2)))
        out = []
        for k in range(D + 1):
            v = c * (4 * k - 1) ** 2
            man, exp = int(v.man), int(v.exp)
            out.append((-1 if v < 0 else 1) * (Fraction(man) * Fraction(2) ** exp))
    return tuple(out)


def sample_deltas(spin: int, lam: int, numax: int, gap: Fraction):
    D = num_poles(numax, spin) + lam
    return [x + gap for x in _sample_points(D)]


def mixed_ising_keys(prec: int, lam: int, numax: int, maxspin: int, ds="0.5181489", de="1.412625"):
    """(spin, Delta, S, P) of every block table a mixed-correlator Ising PMP needs in its continuous sectors
    (model of test/src/main.cpp:97-178): sector "even" (even spins, S=P=0), "odd+" (even spins) and "odd-" (odd spins)
    with the two (S, P) classes of the (e s e s) and (e s s e) orderings."""
    ds, de = Fraction(ds), Fraction(de)
    delta = de - ds
    classes_even = [(Fraction(0), Fraction(0))]
    classes_odd = [(Fraction(0), -delta * delta / 2), (-delta, delta * delta / 2)]
    eps = Fraction(1, 2)
    keys = []

    def gap(l):
        return eps if l == 0 else l + 2 * eps

    for l in range(0, maxspin + 1, 2):
        g = {0: Fraction(6), 2: Fraction(5)}.get(l, gap(l))
        for d in sample_deltas(l, lam, numax, g):
            for S, P in classes_even:
                keys.append((l, d, S, P))
    for l in range(0, maxspin + 1):
        g = Fraction(3) if l == 0 else gap(l)
        for d in sample_deltas(l, lam, numax, g):
            for S, P in classes_odd:
                keys.append((l, d, S, P))
    return keys


def single_ising_keys(prec: int, lam: int, numax: int, maxspin: int):
    """single-correlator Ising (sample/main.cpp:32-59): even spins, S = P = 0"""
    eps = Fraction(1, 2)
    keys = []
    for l in range(0, maxspin + 1, 2):
        g = {0: Fraction(6), 2: Fraction(5)}.get(l, l + 2 * eps)
        for d in sample_deltas(l, lam, numax, g):
            keys.append((l, d, Fraction(0), Fraction(0)))
    return keys


def sweep_keys(prec: int, lam: int, numax: int, maxspin: int, n_tables: int):
    """The scan workload of BASELINE config 5: the mixed-Ising key set repeated over a grid of (Delta_sigma,
    Delta_epsilon) around (0.5181489, 1.412625) in steps of 1e-4 until n_tables keys exist."""
    keys = []
    i = 0
    while len(keys) < n_tables:
        gs, ge = divmod(i, 13)
        ds = Fraction("0.5181489") + Fraction(gs - 6, 10000)
        de = Fraction("1.412625") + Fraction(ge - 6, 10000)
        keys += mixed_ising_keys(prec, lam, numax, maxspin, ds, de)
        i += 1
    return keys[:n_tables]


def to_records(keys, prec: int):
    """keys -> (spin[n] uint32, delta, S, P as (n, W) uint32 record arrays); conversions memoised"""
    cache = {}

    def rec(x):
        r = cache.get(x)
        if r is None:
            r = cache[x] = from_fraction(x, prec)
        return r

    spin = np.array([k[0] for k in keys], dtype=np.uint32)
    delta = np.stack([rec(k[1]) for k in keys])
    S = np.stack([rec(k[2]) for k in keys])
    P = np.stack([rec(k[3]) for k in keys])
    return spin, delta, S, P


def work_per_table(n_max: int, lam: int, limbs: int, spin: int):
    """Algorithmic limb-MACs of ONE table, counting the reference algorithm's operations:
    multiply / fma = L^2, division = 2 L^2, pow = 64 L^2.  Returned per kernel stage and in total."""
    L2 = limbs * limbs
    K, A = (6, 170) if spin == 0 else (8, 720)
    E = {11: 155, 13: 222, 19: 495, 27: 1027}.get(lam)
    if E is None:  # loop-bound count of the transverse recursion's real multiplies
        E = sum((lam - 2 * n + 1) * (5 if n >= 2 else 3) + (lam - 2 * n) for n in range(1, lam // 2 + 1))
    divs_t = sum(lam - 2 * n + 1 for n in range(1, lam // 2 + 1))
    rec = A * L2
    series = ((K - 1) * n_max - (K - 1) * (K - 2) // 2) * L2 + n_max * 2 * L2
    deriv = ((2 * lam + 2) * (n_max + 1) + (lam + 1)) * L2 + lam * 2 * L2 + (lam + 2) * 64 * L2
    finish = ((lam + 1) ** 2 + E) * L2 + divs_t * 2 * L2
    return dict(rec=rec, series=series, deriv=deriv, finish=finish, total=rec + series + deriv + finish)
