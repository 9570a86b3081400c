// qboot_b200 -- exp / pow on the fixed-limb float (host + device).
//
// The reference calls mpfr_pow / mpfr_ui_pow on the table path (src/hor_recursion.cpp:43 `pow(4, Delta)`,
// include/qboot/algebra/real_function.hpp:262 `pow(x, pow_)`), which are correctly rounded.  Here x^y is evaluated as
// exp(y * ln x) in an extended format of E = L + QB_EXT_LIMBS limbs (ln x and ln 2 are supplied at that width) and
// rounded once to `prec` bits.  The extended result carries an error below 2^46 units of ITS last place, so the final
// rounding equals the correct rounding unless the exact value lies within that distance of a rounding boundary (a
// midpoint between two prec-bit numbers).  That case is DETECTED (ziv_safe: the discarded bits read 1000...0x or
// 0111...1x down to bit QB_ZIV_SLACK) and the value is recomputed at E = L + QB_EXT2_LIMBS limbs -- Ziv's strategy,
// which is what makes mpfr_pow correctly rounded; x^y with rational x, y is never exactly a midpoint unless it is
// exactly representable, so the second level settles everything but a 2^-140 event, which is reported, not hidden.
#pragma once

#include "fastops.cuh"
#include "mpf.cuh"

#ifndef QB_EXT_LIMBS
#define QB_EXT_LIMBS 3
#endif
#ifndef QB_EXT2_LIMBS
#define QB_EXT2_LIMBS 6  // second Ziv level
#endif
// error bound of an extended result, in bits of its own last place, with margin (analysis: < 2^46)
#define QB_ZIV_SLACK 56

namespace qb
{
	// widen / narrow ---------------------------------------------------------------------------------------------------
	template <int E, int L>
	QB_HD QB_INLINE void widen(mpf<E>& r, const mpf<L>& a)
	{
		static_assert(E >= L, "widen");
		r.sk = a.sk;
		r.exp = a.exp;
		for (int i = 0; i < E - L; ++i) r.m[i] = 0u;
		for (int i = 0; i < L; ++i) r.m[E - L + i] = a.m[i];
	}
	template <int L, int E>
	QB_HD QB_INLINE void narrow(mpf<L>& r, const mpf<E>& a, int prec)
	{
		if (kind(a) != K_REG)
		{
			set_zero(r, a.sk);
			r.sk = a.sk;
			return;
		}
		round_to<L, E>(r, a.sk, a.exp, a.m, 0u, prec);
	}
	// Can the rounding of `a` (E limbs, true value within 2^slack units of a's last place) to prec bits be trusted?
	// Bit G - 1 of the mantissa (G = 32 E - prec discarded bits) is the round bit; the rounding is in doubt iff the
	// bits below it, down to bit `slack`, all differ from it: 1000...0x / 0111...1x, a value next to a midpoint.
	template <int E>
	QB_HD QB_INLINE bool ziv_safe(const mpf<E>& a, int prec, int slack)
	{
		if (kind(a) != K_REG) return true;
		const int G = 32 * E - prec;
		if (slack >= G - 1) return false;
		const uint32_t rb = (a.m[(G - 1) >> 5] >> ((G - 1) & 31)) & 1u;
		for (int i = G - 2; i >= slack; --i)
			if (((a.m[i >> 5] >> (i & 31)) & 1u) == rb) return true;
		return false;
	}
	// top 53 bits as a double (truncated); zero/regular only
	template <int E>
	QB_HD double to_double(const mpf<E>& a)
	{
		if (kind(a) != K_REG) return 0.0;
		uint64_t hi = (uint64_t(a.m[E - 1]) << 32) | (E >= 2 ? a.m[E - 2] : 0u);
		double d = double(hi >> 11) * (1.0 / 9007199254740992.0);  // [0.5, 1)
		// scale by 2^exp without libm dependencies on device
		int32_t e = a.exp;
		if (e > 1000) e = 1000;
		if (e < -1000) e = -1000;
		double s = 1.0;
		double base = e >= 0 ? 2.0 : 0.5;
		uint32_t n = uint32_t(e >= 0 ? e : -e);
		while (n)
		{
			if (n & 1u) s *= base;
			base *= base;
			n >>= 1;
		}
		d *= s;
		return sign(a) ? -d : d;
	}

	// r = exp(t), all in E limbs at 32E bits.  ln2 = ln 2 at the same width.
	template <int E>
	QB_HD QB_NOINLINE void exp_ext(mpf<E>& r, const mpf<E>& t, const mpf<E>& ln2)
	{
		const int pe = 32 * E;
		if (is_zero(t))
		{
			set_si(r, 1, pe);
			return;
		}
		// k = nearest integer to t / ln 2
		double td = to_double(t);
		double kd = td * 1.4426950408889634;
		int64_t k = int64_t(kd + (kd >= 0 ? 0.5 : -0.5));
		mpf<E> x, u;
		mul_si(u, ln2, k, pe);
		sub(x, t, u, pe);  // |x| <~ 0.35 (+ slack)
		const int s = 24;
		if (kind(x) == K_REG) x.exp -= s;
		// Taylor: sum = 1 + x + x^2/2! + ...
		mpf<E> sum, term;
		set_si(sum, 1, pe);
		copy(term, x);
		for (int n = 1; n < 4 * E + 16; ++n)
		{
			if (is_zero(term)) break;
			if (term.exp < -pe - 4) break;
			fast::add(sum, sum, term, pe);
			fast::mul(u, term, x, pe);
			div_si(term, u, n + 1, pe);
		}
		for (int i = 0; i < s; ++i)
		{
			fast::mul(u, sum, sum, pe);
			copy(sum, u);
		}
		copy(r, sum);
		r.exp += int32_t(k);
	}

	// r = RN_prec(exp(y * lnx)),  lnx and ln2 given in E = L + X limbs.  Returns false when the rounding is in doubt at
	// this width (see ziv_safe): the caller repeats the call with the wider constants.
	template <int L, int X = QB_EXT_LIMBS>
	// r, y thread-local; lnx, ln2 anywhere
	QB_HD QB_NOINLINE bool pow_from_log(mpf<L>& r, const mpf<L + X>& lnx, const mpf<L>& y, const mpf<L + X>& ln2, int prec,
	                                    int slack = QB_ZIV_SLACK)
	{
		constexpr int E = L + X;
		// lnx / ln2 may live in global memory: stage them into thread-local records once, they are reused below
		mpf<E> ye, t, v, lx, l2;
		copy(lx, lnx);
		copy(l2, ln2);
		widen(ye, y);
		fast::mul(t, ye, lx, 32 * E);
		exp_ext(v, t, l2);
		narrow(r, v, prec);
		return ziv_safe(v, prec, slack);
	}

	// Ziv's loop for x^y: first level at L + QB_EXT_LIMBS limbs (lnx1, ln2_1), second at L + QB_EXT2_LIMBS (lnx2, ln2_2).
	// counters (may be null): [0] += second-level evaluations, [1] += values still in doubt after it (reported by the engine
	// as QB_ERR_ROUNDING: the result is then correctly rounded only with probability 1 - 2^-140).
	QB_HD QB_INLINE void ziv_count(uint32_t* counters, int which)
	{
		if (!counters) return;
#if defined(__CUDA_ARCH__)
		atomicAdd(counters + which, 1u);
#else
		counters[which] += 1u;
#endif
	}
	template <int L>
	QB_HD QB_NOINLINE void pow_ziv(mpf<L>& r, const uint32_t* lnx1, const uint32_t* ln2_1, const uint32_t* lnx2, const uint32_t* ln2_2,
	                               const mpf<L>& y, int prec, int slack, uint32_t* counters)
	{
		constexpr int E1 = L + QB_EXT_LIMBS, E2 = L + QB_EXT2_LIMBS;
		if (pow_from_log<L, QB_EXT_LIMBS>(r, *reinterpret_cast<const mpf<E1>*>(lnx1), y, *reinterpret_cast<const mpf<E1>*>(ln2_1), prec, slack))
			return;
		ziv_count(counters, 0);
		if (pow_from_log<L, QB_EXT2_LIMBS>(r, *reinterpret_cast<const mpf<E2>*>(lnx2), y, *reinterpret_cast<const mpf<E2>*>(ln2_2), prec))
			return;
		ziv_count(counters, 1);
	}

#if !defined(__CUDACC__)
	// test harness hook (tests/hostemu): consts = [ln2 (E+2 words), lnx (E+2 words)]; p = Ziv slack override (0: default).
	// A rounding in doubt at this width is reported by returning false: the harness then has no answer (the engine
	// recomputes with the wider constants, see k_pow_fix).
	template <int L>
	bool hostemu_math(int prec, int op, const mpf<L>& x, const mpf<L>& y, int64_t p, const uint32_t* consts, mpf<L>& r)
	{
		constexpr int E = L + QB_EXT_LIMBS;
		if (!consts) return false;
		const mpf<E>* c = reinterpret_cast<const mpf<E>*>(consts);
		const int slack = p > 0 ? int(p) : QB_ZIV_SLACK;
		if (op == 13) return pow_from_log<L>(r, c[1], y, c[0], prec, slack);  // x^y with ln x = c[1]
		if (op == 14) return pow_from_log<L>(r, c[1], x, c[0], prec, slack);  // p^x with ln p = c[1]
		if (op == 15)
		{
			mpf<E> xe, v;
			widen(xe, x);
			exp_ext(v, xe, c[0]);
			narrow(r, v, prec);
			return true;
		}
		return false;
	}
#endif
}  // namespace qb
