// qboot_b200 -- polynomial-matrix-program assembly on the device: building blocks (host + device) of kernels_g5.cuh.
//
//   pmp_entry       one entry M[p][n] of a constraint: sum over the terms of an equation of coeff * F-block entry
//                   (BootstrapEquation::make_cont_mat / make_disc_mat, src/bootstrap_equation.cpp:380-461)
//   pmp_pack_entry  one entry of d_B / d_c: negate, re-index to free variables, eliminate the pivoted variables with fma
//                   (PolynomialProgram::create_input, src/polynomial_program.cpp:194-221)
//   format_decimal  one number as SDPB text: mpfr_get_str's digits (round to nearest, ties to even) laid out by the
//                   reference's default-float rules (include/qboot/mp/real.hpp:938-956,987-1030; src/sdpb_input.cpp:32-35)
//
// As everywhere in this engine the operations and their order are the reference's, each rounded once at `prec` bits.
#pragma once

#include <math.h>

#include "block.cuh"
#include "pmp_types.cuh"

namespace qb
{
	// (M, N) of the i-th kept entry of a block of parity sym, N-major / M-minor (complex_function.hpp:72-81)
	QB_HD QB_INLINE void cf_decode_sym(int lambda, int sym, int i, int& M, int& N)
	{
		N = 0;
		for (;; ++N)
		{
			int top = lambda - 2 * N - sym;
			int row = top < 0 ? 0 : top / 2 + 1;
			if (i < row) break;
			i -= row;
		}
		M = 2 * i + sym;
	}

	// M[p][n] for variable n = eq_ofs[eq] + i, position rc, sample k.  `terms` are the block's terms in insertion order;
	// every term of this (eq, rc) contributes coef * F_term[i], added in order onto +0.
	// g: tables (dim_M records each); v: v-power tables; coef: coefficient records.  out thread-local.
	template <int L>
	QB_HD QB_NOINLINE void pmp_entry(mpf<L>& out, const PmpTerm* terms, int n_terms, int eq, int rc, int k, int sym, int i,
	                                 const int32_t* tab, const mpf<L>* g, const mpf<L>* v, const mpf<L>* coef, int lambda, int prec)
	{
		const int dimM = cf_dim(lambda);
		int M, N;
		cf_decode_sym(lambda, sym, i, M, N);
		mpf<L> acc, f, c, t;
		set_zero(acc);
		for (int j = 0; j < n_terms; ++j)
		{
			const PmpTerm tm = terms[j];
			if (tm.eq != eq || tm.rc != rc) continue;
			const mpf<L>* gt = g + size_t(tab[tm.tab0 + k]) * dimM;
			fblock_entry(f, v + size_t(tm.d) * dimM, 1, gt, 1, lambda, M, N, prec);
			if (tm.coef >= 0)
			{
				copy(c, coef[tm.coef]);
				fast::mul(t, c, f, prec);
				fast::add(acc, acc, t, prec);
			}
			else
				fast::add(acc, acc, f, prec);
		}
		copy(out, acc);
	}

	// d_B[p][m] = -M[p][free_m], then for each eliminated variable e: fma(eqn[e][free_m], M[p][lead_e], .)
	// column m == n_free gives d_c[p] = -target[p] (target may be null = +0), then fma(eqn_target[e], M[p][lead_e], .)
	// Mp: row p of M (N records); eqn: n_elim x N records; out thread-local
	template <int L>
	QB_HD QB_NOINLINE void pmp_pack_entry(mpf<L>& out, const mpf<L>* Mp, int m, int n_free, const int32_t* free_idx, int n_elim,
	                                      const int32_t* lead, const mpf<L>* eqn, const mpf<L>* eqn_target, const mpf<L>* target_p,
	                                      int N, int prec)
	{
		mpf<L> acc, a, t;
		if (m < n_free)
			copy(acc, Mp[free_idx[m]]);
		else if (target_p)
			copy(acc, *target_p);
		else
			set_zero(acc);
		neg(acc);
		for (int e = 0; e < n_elim; ++e)
		{
			copy(t, Mp[lead[e]]);
			if (m < n_free)
				copy(a, eqn[size_t(e) * N + free_idx[m]]);
			else
				copy(a, eqn_target[e]);
			fast::fma(acc, a, t, acc, prec);
		}
		copy(out, acc);
	}

	// out = RN(eval(c[0 .. deg], x) * s): Horner with fused multiply-adds, exactly Polynomial::eval with a real argument
	// (include/qboot/algebra/polynomial.hpp:66-83), then one product -- an entry q_m(x_k) sqrt(s_k) of the bilinear bases at
	// the sample points (src/polynomial_program.cpp:224-229).  deg < 0: the zero polynomial.  c, x, s anywhere; out thread-local
	template <int L>
	QB_HD QB_NOINLINE void poly_eval_scaled(mpf<L>& out, const mpf<L>* c, int deg, const mpf<L>& x, const mpf<L>& s, int prec)
	{
		mpf<L> acc, xx, cc, ss;
		if (deg < 0)
			set_zero(acc);
		else
		{
			copy(acc, c[deg]);
			copy(xx, x);
			for (int i = deg - 1; i >= 0; --i)
			{
				copy(cc, c[i]);
				fast::fma(acc, acc, xx, cc, prec);
			}
		}
		copy(ss, s);
		fast::mul(out, acc, ss, prec);
	}

	// ---- decimal text ---------------------------------------------------------------------------------------------------
	QB_HD QB_INLINE uint32_t pow5_small(int b)
	{
		uint32_t r = 1;
		for (int i = 0; i < b; ++i) r *= 5u;
		return r;
	}
	// floor(log10 |x|) + 1 up to +-1 (the caller corrects): log10(2) * (exp - 1 + log2(top bits))
	template <int L>
	QB_HD QB_INLINE int dec_exponent_guess(const mpf<L>& x)
	{
		// top 32 bits as m in [1, 2); a guess that is off by one next to a power of ten is corrected by the caller
		double m = double(x.m[L - 1]) * (1.0 / 2147483648.0);
		double v = (double(x.exp) - 1.0 + log2(m)) * 0.30102999566398120;
		int f = int(v);
		if (double(f) > v) --f;
		return f + 1;
	}
	// out: sign, digits, point / exponent, '\n'.  Returns the length, or a negative code when the number is outside the range
	// handled here (huge or tiny exponents; the host formats those): -1 digit count, -2 inf/nan, -3 exponent search did not
	// settle, -4 |x| >= 10^ndigits or the power of five is beyond the table, -5 / -6 / -7 the scaled integer does not fit.
	template <int L>
	QB_HD QB_NOINLINE int format_decimal(char* out, const mpf<L>& x, const DecTables& tb)
	{
		constexpr int NW = L + 2;            // limbs of N (the digit integer) with headroom
		constexpr int PW = L + QB_DEC_MAXA;  // limbs of M * 5^k
		constexpr int MAXD = 9 * ((L * 32 * 302 / 1000 + 3 + 8) / 9 + 1);
		const int n = tb.ndigits;
		int len = 0;
		if (n < 1 || n > MAXD) return -1;
		if (kind(x) == K_ZERO)
		{
			if (sign(x)) out[len++] = '-';
			out[len++] = '0';
			out[len++] = '\n';
			return len;
		}
		if (kind(x) != K_REG) return -2;
		int d = dec_exponent_guess(x);
		uint32_t Nn[NW];
		for (int attempt = 0;; ++attempt)
		{
			if (attempt == 3) return -3;
			const int k = n - d;
			if (k < 0 || k / 13 >= QB_DEC_MAXA) return -4;
			// N = RN(M * 5^k * 2^k / 2^(32 L - exp)) = RN((M * 5^k) >> s)
			const int64_t s = int64_t(32) * L - int64_t(x.exp) - k;
			if (s < 1) return -5;
			const int a = k / 13, b = k % 13;
			const uint32_t* P = tb.p13 + tb.pofs[a];
			const int pl = tb.pofs[a + 1] - tb.pofs[a];
			if (L + pl + 1 > PW) return -6;
			uint32_t prod[PW + 1];
			const int tl = L + pl + 1;  // limbs in use
			for (int i = 0; i < tl; ++i) prod[i] = 0u;
			// schoolbook M * P
			QB_ROLL
			for (int i = 0; i < pl; ++i)
			{
				uint64_t carry = 0;
				const uint32_t pi = P[i];
				QB_ROLL
				for (int j = 0; j < L; ++j)
				{
					uint64_t t = uint64_t(pi) * x.m[j] + prod[i + j] + carry;
					prod[i + j] = uint32_t(t);
					carry = t >> 32;
				}
				prod[i + L] = uint32_t(carry);
			}
			if (b)
			{
				const uint32_t f = pow5_small(b);
				uint64_t carry = 0;
				QB_ROLL
				for (int i = 0; i < tl; ++i)
				{
					uint64_t t = uint64_t(prod[i]) * f + carry;
					prod[i] = uint32_t(t);
					carry = t >> 32;
				}
			}
			// shift right by s with round bit and sticky
			const int q = int(s >> 5), r = int(s & 31);
			if (q >= tl + 1) return -7;
			// bit s - 1 is the round bit, everything below it is sticky
			const int64_t rb = s - 1;
			const int rq = int(rb >> 5), rr = int(rb & 31);
			const bool rnd = rq < tl && ((prod[rq] >> rr) & 1u);
			bool sticky = rq < tl && rr && (prod[rq] & ((1u << rr) - 1u));
			for (int i = 0; i < rq && i < tl; ++i)
				if (prod[i]) sticky = true;
			bool overflow = false;
			for (int i = 0; i < NW; ++i)
			{
				const int lo = q + i;
				uint32_t w0 = lo < tl ? prod[lo] : 0u, w1 = lo + 1 < tl ? prod[lo + 1] : 0u;
				Nn[i] = r ? ((w0 >> r) | (w1 << (32 - r))) : w0;
			}
			for (int i = q + NW; i < tl; ++i)
				if (prod[i] >> (i == q + NW ? r : 0)) overflow = true;
			if (overflow)
			{
				++d;  // far too many digits: the guess was low
				continue;
			}
			// the unrounded quotient decides the exponent: 10^(n-1) <= N < 10^n
			auto cmpw = [&](const uint32_t* B) {
				for (int i = NW - 1; i >= 0; --i)
					if (Nn[i] != B[i]) return Nn[i] < B[i] ? -1 : 1;
				return 0;
			};
			if (cmpw(tb.ten_n) >= 0)
			{
				++d;
				continue;
			}
			if (cmpw(tb.ten_n1) < 0)
			{
				--d;
				continue;
			}
			if (rnd && (sticky || (Nn[0] & 1u)))
			{
				for (int i = 0; i < NW; ++i)
					if (++Nn[i] != 0u) break;
				if (cmpw(tb.ten_n) >= 0)
				{
					for (int i = 0; i < NW; ++i) Nn[i] = tb.ten_n1[i];
					++d;
				}
			}
			break;
		}
		// digits of N, most significant first, into dg[0..n)
		char dg[MAXD];
		{
			int top = NW;
			int pos = n;
			while (pos > 0)
			{
				while (top > 0 && Nn[top - 1] == 0u) --top;
				uint64_t rem = 0;
				for (int i = top - 1; i >= 0; --i)
				{
					uint64_t cur = (rem << 32) | Nn[i];
					Nn[i] = uint32_t(cur / 1000000000u);
					rem = cur % 1000000000u;
				}
				uint32_t chunk = uint32_t(rem);
				for (int j = 0; j < 9 && pos > 0; ++j)
				{
					dg[--pos] = char('0' + chunk % 10u);
					chunk /= 10u;
				}
			}
		}
		if (sign(x)) out[len++] = '-';
		int nd = n;  // significant digits kept after stripping zeros
		while (nd > 1 && dg[nd - 1] == '0') --nd;
		if (-3 <= d && d <= n)
		{
			// fixed notation
			if (d <= 0)
			{
				out[len++] = '0';
				out[len++] = '.';
				for (int i = 0; i < -d; ++i) out[len++] = '0';
				for (int i = 0; i < nd; ++i) out[len++] = dg[i];
			}
			else
			{
				for (int i = 0; i < d; ++i) out[len++] = i < n ? dg[i] : '0';
				if (nd > d)
				{
					out[len++] = '.';
					for (int i = d; i < nd; ++i) out[len++] = dg[i];
				}
			}
		}
		else
		{
			out[len++] = dg[0];
			if (nd > 1)
			{
				out[len++] = '.';
				for (int i = 1; i < nd; ++i) out[len++] = dg[i];
			}
			int ex = d - 1;
			out[len++] = 'e';
			out[len++] = ex >= 0 ? '+' : '-';
			if (ex < 0) ex = -ex;
			char eb[12];
			int el = 0;
			do
			{
				eb[el++] = char('0' + ex % 10);
				ex /= 10;
			} while (ex);
			if (el < 2) eb[el++] = '0';
			while (el) out[len++] = eb[--el];
		}
		out[len++] = '\n';
		return len;
	}
}  // namespace qb
