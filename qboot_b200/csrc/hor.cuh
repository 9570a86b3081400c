// qboot_b200 -- Hogervorst-Osborn-Rychkov series recurrence on the fixed-limb float (host + device).
//
//   rec_coeffs   : coefficients of  sum_i p_i(n) b[n-i] = 0   (reference: _get_rec_coeffs, include/qboot/hor_formula.hpp:34-39,
//                  src/hor_formula.cpp:8-296 spin 0, :298-2055 spin > 0)
//   series       : b[n] = -(sum_{i=1}^{min(n,K-1)} p_i(n) b[n-i]) / p_0(n)   (src/hor_recursion.cpp:31-37)
//
// The middle polynomials p_1..p_{K-2} come from the integer tensors in hor_table.inc (oracle/derive_hor_table.py);
// each coefficient is evaluated monomial by monomial in lexicographic order, every product and every partial sum
// rounded at `prec` bits, which reproduces the reference's MPFR evaluation bit for bit.
//
// Operands may live in registers, thread-local, shared or global memory: all arithmetic is force-inlined (mpf.cuh), so
// every limb access compiles to the load of the space the operand really lives in.  The register fast paths of
// fastops.cuh are used wherever an operation sits on a hot loop; the generic routines of mpf.cuh elsewhere.
#pragma once

#include "fastops.cuh"
#include "mpf.cuh"

namespace qb
{
	struct HorMono
	{
		int32_t c;
		uint8_t eD, eS, eP, el, ee;
	};
	struct HorPoly
	{
		int32_t first, count, konst, flag;
	};
	// epsilon = (dim - 2) / 2 as an exact rational
	struct Rational
	{
		int64_t num;
		uint32_t den;
	};

	// value of one monomial |c| * Delta^a S^b P^c spin^d eps^e (sign handled by the caller unless `signed_c`);
	// t, D, S, P are thread-local
	template <int L>
	QB_HD QB_NOINLINE void hor_monomial(mpf<L>& t, HorMono mo, bool signed_c, const mpf<L>& D, const mpf<L>& S,
	                                    const mpf<L>& P, uint32_t spin, Rational eps, int prec)
	{
		int64_t c = mo.c;
		if (!signed_c && c < 0) c = -c;
		int nD = mo.eD, nS = mo.eS, nP = mo.eP, nl = mo.el, ne = mo.ee;
		// leading factor
		if (nD)
		{
			if (c == 1)
				copy(t, D);
			else
				fast::mul_small(t, D, c, prec);
			--nD;
		}
		else if (nS)
		{
			if (c == 1)
				copy(t, S);
			else
				fast::mul_small(t, S, c, prec);
			--nS;
		}
		else if (nP)
		{
			if (c == 1)
				copy(t, P);
			else
				fast::mul_small(t, P, c, prec);
			--nP;
		}
		else if (nl)
		{
			// integer arithmetic: exact
			int64_t v = c;
			for (; nl > 0; --nl) v *= int64_t(spin);
			set_si(t, v, prec);
		}
		else
		{
			set_q(t, c * eps.num, eps.den, prec);
			--ne;
		}
		for (; nD > 0; --nD) fast::mul(t, t, D, prec);
		for (; nS > 0; --nS) fast::mul(t, t, S, prec);
		for (; nP > 0; --nP) fast::mul(t, t, P, prec);
		for (; nl > 0; --nl) fast::mul_small(t, t, int64_t(spin), prec);
		for (; ne > 0; --ne)
		{
			// eps = +-1 / 2^k (dimensions 3, 2.5, ...): the rounded product is an exact exponent shift
			if ((eps.num == 1 || eps.num == -1) && (eps.den & (eps.den - 1)) == 0 && kind(t) == K_REG)
			{
				int sh = 0;
				while ((eps.den >> sh) > 1u) ++sh;
				t.exp -= sh;
				if (eps.num < 0) t.sk ^= SIGN_BIT;
			}
			else
				mul_q(t, t, eps.num, eps.den, prec);
		}
	}

	// one coefficient a[i][d]; r, D, S, P thread-local, the tables anywhere
	template <int L>
	QB_HD QB_NOINLINE void hor_coefficient(mpf<L>& r, HorPoly po, const HorMono* monos, const mpf<L>& D, const mpf<L>& S,
	                                       const mpf<L>& P, uint32_t spin, Rational eps, int prec)
	{
		mpf<L> acc, t;
		if (po.count == 0)
			set_si(acc, po.konst, prec);
		else
		{
			hor_monomial(acc, monos[po.first], true, D, S, P, spin, eps, prec);
			for (int k = 1; k < po.count; ++k)
			{
				HorMono mo = monos[po.first + k];
				hor_monomial(t, mo, false, D, S, P, spin, eps, prec);
				fast::add(acc, acc, t, prec, mo.c < 0);
			}
			fast::add_small(acc, acc, po.konst, prec);
		}
		if (po.flag == 1)
		{
			// times (3 - 2 Delta + 2 eps)
			mpf<L> lin, e2;
			mul_si(lin, D, -2, prec);
			set_q(e2, 2 * eps.num, eps.den, prec);
			add(lin, lin, e2, prec);
			add_si(lin, lin, 3, prec);
			mul(r, lin, acc, prec);
		}
		else
			copy(r, acc);
	}

	// c[0..m] (degree m) times (n + a)  ->  c[0..m+1]     (reference: Polynomial::_mul_linear, src/algebra/polynomial.cpp:76-83)
	// c is a thread-local array
	template <int L>
	QB_HD QB_NOINLINE void mul_linear(mpf<L>* c, int m, const mpf<L>& a, int prec)
	{
		// in place, top down: new[m+1] = c[m]; new[i] = c[i] * a + c[i-1]; new[0] = c[0] * a
		copy(c[m + 1], c[m]);
		for (int i = m; i >= 1; --i) fast::fma(c[i], c[i], a, c[i - 1], prec);
		fast::mul(c[0], c[0], a, prec);
	}

	// p_0 (first == true) or p_{K-1}: products of linear factors (n + a), built by repeated mul_linear exactly as the
	// reference does (src/hor_formula.cpp:16-19,291-294 spin 0; :306-311,2049-2054 spin > 0).  p[0..deg] thread-local.
	template <int L>
	QB_HD QB_NOINLINE void rec_coeffs_end(mpf<L>* p, const mpf<L>& D, uint32_t spin, Rational eps, int prec, bool first)
	{
		mpf<L> a;
		if (spin == 0)
		{
			if (first)
			{
				// p_0 = n (n + Delta - 1) (n + 2 Delta - 2 eps - 2)
				set_zero(p[0]);
				set_si(p[1], 1, prec);
				add_si(a, D, -1, prec);
				mul_linear(p, 1, a, prec);
				mul_si(a, D, 2, prec);
				sub_q(a, a, 2 * eps.num + 2 * int64_t(eps.den), eps.den, prec);
				mul_linear(p, 2, a, prec);
			}
			else
			{
				// p_5 = (n + Delta - 4) (n + 2 Delta - 5) (n + 2 eps - 3)
				add_si(p[0], D, -4, prec);
				set_si(p[1], 1, prec);
				mul_si(a, D, 2, prec);
				add_si(a, a, -5, prec);
				mul_linear(p, 1, a, prec);
				set_q(a, 2 * eps.num - 3 * int64_t(eps.den), eps.den, prec);
				mul_linear(p, 2, a, prec);
			}
		}
		else if (first)
		{
			// p_0 = -n (n + Delta + spin - 1) (n + 2 Delta - 2 eps - 2) (n + Delta - 2 eps - spin - 1)
			set_zero(p[0]);
			set_si(p[1], -1, prec);
			add_si(a, D, int64_t(spin) - 1, prec);
			mul_linear(p, 1, a, prec);
			mul_si(a, D, 2, prec);
			sub_q(a, a, 2 * eps.num, eps.den, prec);
			add_si(a, a, -2, prec);
			mul_linear(p, 2, a, prec);
			sub_q(a, D, 2 * eps.num, eps.den, prec);
			add_si(a, a, -int64_t(spin), prec);
			add_si(a, a, -1, prec);
			mul_linear(p, 3, a, prec);
		}
		else
		{
			// p_7 = (n + 2 Delta - 7) (n + 2 eps - 5) (n + Delta - spin - 6) (n + Delta + 2 eps + spin - 6)
			mul_si(p[0], D, 2, prec);
			add_si(p[0], p[0], -7, prec);
			set_si(p[1], 1, prec);
			set_q(a, 2 * eps.num - 5 * int64_t(eps.den), eps.den, prec);
			mul_linear(p, 1, a, prec);
			add_si(a, D, -(int64_t(spin) + 6), prec);
			mul_linear(p, 2, a, prec);
			add_q(a, D, 2 * eps.num + (int64_t(spin) - 6) * int64_t(eps.den), eps.den, prec);
			mul_linear(p, 3, a, prec);
		}
	}

	// All recurrence coefficients of one table into out[i * (deg + 1) + d] (out anywhere), i < K,
	// K = 6 / deg = 3 for spin 0, else 8 / 4.  Serial convenience used by the host harness.
	template <int L>
	QB_HD void rec_coeffs(mpf<L>* out, const mpf<L>& Din, uint32_t spin, const mpf<L>& Sin, const mpf<L>& Pin, Rational eps,
	                      int prec, const HorPoly* polys, const HorMono* monos)
	{
		const int K = spin == 0 ? 6 : 8, deg = spin == 0 ? 3 : 4, st = deg + 1;
		mpf<L> D, S, P, r, pe[5];
		copy(D, Din);
		copy(S, Sin);
		copy(P, Pin);
		for (int i = 1; i < K - 1; ++i)
			for (int d = 0; d <= deg; ++d)
			{
				hor_coefficient(r, polys[(i - 1) * st + d], monos, D, S, P, spin, eps, prec);
				copy(out[i * st + d], r);
			}
		rec_coeffs_end(pe, D, spin, eps, prec, true);
		for (int d = 0; d <= deg; ++d) copy(out[d], pe[d]);
		rec_coeffs_end(pe, D, spin, eps, prec, false);
		for (int d = 0; d <= deg; ++d) copy(out[(K - 1) * st + d], pe[d]);
	}

	// p(n) by Horner with the integer n: s = s * n, s = s + c   (include/qboot/algebra/polynomial.hpp:66-83); c thread-local
	template <int L>
	QB_HD QB_NOINLINE void poly_eval_ui(mpf<L>& s, const mpf<L>* c, int deg, uint32_t n, int prec)
	{
		copy(s, c[deg]);
		for (int d = deg - 1; d >= 0; --d)
		{
			fast::mul_small(s, s, int64_t(n), prec);
			fast::add(s, s, c[d], prec);
		}
	}

	// b[0] .. b[n_max] of one table, b[0] given (b anywhere, addressed as b[n * stride]); `coef` from rec_coeffs (anywhere).
	// Coefficients and the last seven series terms are kept in thread-local storage.
	template <int L>
	QB_HD void series(mpf<L>* b, int64_t stride, uint32_t n_max, const mpf<L>* coef, uint32_t spin, int prec)
	{
		const int K = spin == 0 ? 6 : 8, deg = spin == 0 ? 3 : 4, st = deg + 1;
		mpf<L> lc[40], hist[8], sum, pv, bn;
		for (int i = 0; i < K * st; ++i) copy(lc[i], coef[i]);
		copy(hist[0], b[0]);
		for (uint32_t n = 1; n <= n_max; ++n)
		{
			set_zero(sum);
			for (int i = 1; i < K; ++i)
				if (uint32_t(i) <= n)
				{
					poly_eval_ui(pv, lc + i * st, deg, n, prec);
					fast::fma(sum, pv, hist[(n - uint32_t(i)) & 7u], sum, prec);
				}
			neg(sum);
			poly_eval_ui(pv, lc, deg, n, prec);
			div(bn, sum, pv, prec);
			copy(hist[n & 7u], bn);
			copy(b[int64_t(n) * stride], bn);
		}
	}
}  // namespace qb
