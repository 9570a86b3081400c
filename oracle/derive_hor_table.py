#!/usr/bin/env python
"""Derive the coefficient polynomials of the Hogervorst-Osborn-Rychkov series recurrence as exact integer tensors.

The recurrence  sum_i p_i(n) b[n-i] = 0  (reference: include/qboot/hor_formula.hpp:12-22) has coefficients
p_i(n) = sum_d a[i][d] n^d where every a[i][d] is a polynomial with integer coefficients in (Delta, S, P, spin,
epsilon).  These are mathematical facts (the power-series form of the Casimir ODE on the diagonal), so instead of
transcribing anything they are RECOVERED here by exact interpolation: the reference build in oracle/_ref is probed
as a black box at small integer points (where 256-bit arithmetic is exact) and the tensor-product Vandermonde systems
are solved over the rationals.  Output (committed):

  qboot_b200/csrc/hor_table.inc   C tables used by the CUDA kernels
  oracle/hor_table.json           the same data for the Python oracle

Monomials are emitted in the canonical order the engine evaluates them in: graded by nothing, pure lexicographic on
exponents with variable order (Delta, S, P, spin, epsilon), highest first; the constant term last.  a[1][0] contains
the factor (3 - 2 Delta + 2 epsilon) in both families and is stored factored (flag 1), which is how it is cheapest to
evaluate.  Run from the repo root in the build container (needs oracle/_ref).
"""
import itertools
import json
import os
import sys
from fractions import Fraction

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref  # noqa: E402

PREC = 512
NPTS = 6  # per-variable degree <= 5
VARS = ["Delta", "S", "P", "spin", "eps"]


def num(v):
    return ref.from_int_exp(0 if v >= 0 else 1, abs(v), 0, PREC)


def probe(spin_family):
    """values[iD, iS, iP, (il,) ie] -> array (K, deg+1) of ints"""
    pts = list(range(NPTS))
    lpts = [0] if spin_family == 0 else list(range(1, NPTS + 1))
    K, deg = (6, 3) if spin_family == 0 else (8, 4)
    shape = (NPTS, NPTS, NPTS, len(lpts), NPTS, K, deg + 1)
    vals = np.empty(shape, dtype=object)
    for iD, iS, iP, il, ie in itertools.product(range(NPTS), range(NPTS), range(NPTS), range(len(lpts)), range(NPTS)):
        out = ref.rec_coeffs(PREC, str(2 * pts[ie] + 2), num(pts[iD]), lpts[il], num(pts[iS]), num(pts[iP]))
        for i in range(K):
            for d in range(deg + 1):
                f = ref.to_fraction(out[i, d], PREC)
                assert f.denominator == 1
                vals[iD, iS, iP, il, ie, i, d] = f.numerator
    return vals, pts, lpts


def inv_vandermonde(xs):
    n = len(xs)
    V = [[Fraction(x) ** j for j in range(n)] for x in xs]
    # Gauss-Jordan over Fractions
    A = [row[:] + [Fraction(int(i == j)) for j in range(n)] for i, row in enumerate(V)]
    for c in range(n):
        p = next(r for r in range(c, n) if A[r][c] != 0)
        A[c], A[p] = A[p], A[c]
        inv = 1 / A[c][c]
        A[c] = [v * inv for v in A[c]]
        for r in range(n):
            if r != c and A[r][c] != 0:
                f = A[r][c]
                A[r] = [a - f * b for a, b in zip(A[r], A[c])]
    return [row[n:] for row in A]


def interpolate(vals, axes_pts):
    """tensor-product interpolation along the first len(axes_pts) axes; returns coefficient tensor (object Fractions)"""
    T = vals
    for ax, xs in enumerate(axes_pts):
        if len(xs) == 1:
            continue
        Vi = inv_vandermonde(xs)
        T = np.moveaxis(T, ax, 0)
        new = np.empty_like(T)
        for j in range(len(xs)):
            acc = None
            for k in range(len(xs)):
                term = T[k] * Vi[j][k] if Vi[j][k] != 0 else None
                if term is not None:
                    acc = term if acc is None else acc + term
            new[j] = acc if acc is not None else T[0] * 0
        T = np.moveaxis(new, 0, ax)
    return T


def monomials(coeff_tensor):
    """coeff_tensor[eD, eS, eP, el, ee] -> list of (coef, (eD, eS, eP, el, ee)) lex-descending, constant separately"""
    mons = []
    const = 0
    for idx in itertools.product(*[range(s) for s in coeff_tensor.shape]):
        c = coeff_tensor[idx]
        c = Fraction(c)
        if c == 0:
            continue
        assert c.denominator == 1, (idx, c)
        if sum(idx) == 0:
            const = int(c)
        else:
            mons.append((int(c), tuple(int(x) for x in idx)))
    mons.sort(key=lambda m: m[1], reverse=True)
    return mons, const


def divide_by_factor(coeff_tensor):
    """exact division of the polynomial by (3 - 2 Delta + 2 eps) using sympy"""
    import sympy as sp

    D, S, P, l, e = sp.symbols("D S P l e")
    syms = (D, S, P, l, e)
    poly = 0
    for idx in itertools.product(*[range(s) for s in coeff_tensor.shape]):
        c = Fraction(coeff_tensor[idx])
        if c != 0:
            term = sp.Integer(c.numerator)
            for s_, k in zip(syms, idx):
                term *= s_**k
            poly += term
    q, r = sp.div(sp.Poly(poly, *syms), sp.Poly(3 - 2 * D + 2 * e, *syms))
    assert r.is_zero, "a[1][0] does not contain the expected linear factor"
    shape = coeff_tensor.shape
    out = np.zeros(shape, dtype=object)
    for mon, c in q.terms():
        out[mon] = Fraction(int(c))
    return out


def main():
    families = {}
    for fam in (0, 1):
        vals, pts, lpts = probe(fam)
        T = interpolate(vals, [pts, pts, pts, lpts, pts])
        K, deg = (6, 3) if fam == 0 else (8, 4)
        polys = {}
        for i in range(1, K - 1):
            for d in range(deg + 1):
                ct = T[..., i, d]
                flag = 0
                if i == 1 and d == 0:
                    ct = divide_by_factor(ct)
                    flag = 1
                mons, const = monomials(ct)
                maxdeg = max([sum(m[1]) for m in mons], default=0)
                polys[f"{i},{d}"] = dict(flag=flag, const=const, mons=mons, total_degree=maxdeg)
        families[str(fam)] = dict(K=K, deg=deg, polys=polys)
        nm = sum(len(p["mons"]) for p in polys.values())
        nmul = sum(sum(m[1]) for p in polys.values() for m in p["mons"])
        print(f"family spin{'=0' if fam == 0 else '>0'}: {nm} monomials, {nmul} variable factors", file=sys.stderr)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "oracle", "hor_table.json"), "w") as f:
        json.dump(families, f, separators=(",", ":"))
    # C tables
    lines = ["// GENERATED by oracle/derive_hor_table.py -- do not edit.",
             "// Integer coefficient tensors of the HOR series recurrence, recovered by exact interpolation.",
             "// mono = {coef, eDelta, eS, eP, espin, eeps}; poly = {first mono, #monos, constant term, flag}",
             "// flag 1: the stored polynomial is to be multiplied by (3 - 2 Delta + 2 eps).",
             "#ifndef QB_TABLE_QUAL", "#define QB_TABLE_QUAL static", "#endif", ""]
    for fam in ("0", "1"):
        F = families[fam]
        K, deg = F["K"], F["deg"]
        monos = []
        polyrows = []
        for i in range(1, K - 1):
            for d in range(deg + 1):
                p = F["polys"][f"{i},{d}"]
                polyrows.append((len(monos), len(p["mons"]), p["const"], p["flag"]))
                monos += p["mons"]
        lines.append(f"QB_TABLE_QUAL const qb::HorMono QB_HOR_MONO_{fam}[{max(len(monos), 1)}] = {{")
        for c, e in monos:
            lines.append(f"\t{{{c}, {e[0]}, {e[1]}, {e[2]}, {e[3]}, {e[4]}}},")
        lines.append("};")
        lines.append(f"QB_TABLE_QUAL const qb::HorPoly QB_HOR_POLY_{fam}[{len(polyrows)}] = {{")
        for r in polyrows:
            lines.append(f"\t{{{r[0]}, {r[1]}, {r[2]}, {r[3]}}},")
        lines.append("};")
        lines.append(f"#define QB_HOR_NMONO_{fam} {len(monos)}")
        lines.append("")
    with open(os.path.join(root, "qboot_b200", "csrc", "hor_table.inc"), "w") as f:
        f.write("\n".join(lines))


if __name__ == "__main__":
    main()
