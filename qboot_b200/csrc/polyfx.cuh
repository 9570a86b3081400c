// qboot_b200 -- the recurrence polynomials p_i(n) in FIXED POINT (host + device).
//
// k_series evaluates K polynomials of degree <= 4 at every step n of every job by the reference's Horner scheme with an
// integer argument (Polynomial::eval, include/qboot/algebra/polynomial.hpp:66-83):
//     s <- c_deg;   s <- RN(RN(s * n) + c_d),  d = deg-1 .. 0          (mul_ui and add, each rounded to prec bits)
// In floating point every step pays two normalisations and an alignment by an arbitrary distance (~1 000 instructions,
// half of the kernel).  The same values are produced here in a fixed-point frame chosen once per polynomial:
//   * a value is (sign, V), V an integer of W = L + 2 limbs, meaning 0.V * 2^E; E is an upper bound for every
//     intermediate of every n <= n_lim, so nothing overflows at the top;
//   * V * n and V +- C are EXACT integer operations (no alignment: the coefficients are stored in the same frame);
//   * RN to prec significant bits is emulated on the integer: the bits below position bitlen(V) - prec are rounded away
//     to nearest-even.  A value with at most prec significant bits is left alone -- it is exactly representable.
// Every result therefore equals the reference's bit for bit PROVIDED the coefficients are exactly representable in the
// frame (checked once, fx_from_record) and the top of V stays inside the three highest limbs (checked at every rounding;
// deep cancellations fall outside).  Otherwise the caller evaluates that polynomial in floating point as before.
#pragma once

#include "fastops.cuh"

namespace qb
{
	namespace fx
	{
		// frame exponent: B_deg = e_deg, B_d = max(B_{d+1} + nbits, e_d) + 1 bounds |s_d| < 2^(B_d) for every n < 2^nbits.
		// e[d] = exponent of c_d, have[d] = c_d is a regular number
		QB_HD QB_INLINE int32_t frame_exponent(const int32_t* e, const bool* have, int deg, int nbits)
		{
			bool any = false;
			int32_t B = 0;
			for (int d = deg; d >= 0; --d)
			{
				if (any)
				{
					B += nbits;
					if (have[d] && e[d] > B) B = e[d];
					B += 1;
				}
				else if (have[d])
				{
					B = e[d];
					any = true;
				}
			}
			return B + 1;
		}
		// V (W = L + 2 limbs) = |c| in the frame of exponent E; false when c is not exactly representable there
		template <int L>
		QB_HD QB_INLINE bool from_record(uint32_t* V, uint32_t& sgn, const mpf<L>& c, int32_t E)
		{
			constexpr int W = L + 2;
			sgn = c.sk & SIGN_BIT;
			if (kind(c) == K_ZERO)
			{
#pragma unroll
				for (int j = 0; j < W; ++j) V[j] = 0u;
				return true;
			}
			if (kind(c) != K_REG) return false;
			const int64_t sh = int64_t(E) - int64_t(c.exp);
			if (sh < 0 || sh > 32 * W) return false;
#pragma unroll
			for (int j = 0; j < W; ++j) V[j] = j >= 2 ? c.m[j - 2] : 0u;
			uint32_t lost = 0u;
			fast::shr_frame<W>(V, int(sh >> 5), int(sh & 31), lost);
			return lost == 0u;
		}
		// RN to prec significant bits, nearest-even, on the integer V.  False: the top of V is below the three highest
		// limbs (the rounding position would leave the three lowest) -- the caller falls back.  V == 0 is fine.
		template <int L>
		QB_HD QB_INLINE bool round_sig(uint32_t* V, int prec)
		{
			constexpr int W = L + 2;
			int jt;
			uint32_t top;
			if (V[W - 1])
			{
				top = V[W - 1];
				jt = W - 1;
			}
			else if (V[W - 2])
			{
				top = V[W - 2];
				jt = W - 2;
			}
			else if (V[W - 3])
			{
				top = V[W - 3];
				jt = W - 3;
			}
			else
			{
				uint32_t nz = 0u;
#pragma unroll
				for (int j = 0; j < W - 3; ++j) nz |= V[j];
				return nz == 0u;  // zero needs no rounding; anything else is outside the window
			}
			const int cut = 32 * jt + 32 - clz32(top) - prec;  // bits to drop
			if (cut <= 0) return true;
			// 1 <= cut <= 32 W - prec <= 95: everything happens in the three lowest limbs
			const int rbp = cut - 1;
			uint32_t rb = 0u, sticky = 0u, lsb = 0u;
#pragma unroll
			for (int j = 0; j < 3; ++j)
			{
				const int lo = 32 * j;
				// round bit and the bits below it
				if (rbp >= lo && rbp < lo + 32) rb = (V[j] >> (rbp - lo)) & 1u;
				const uint32_t below = rbp >= lo + 32 ? 0xFFFFFFFFu : (rbp > lo ? ((1u << (rbp - lo)) - 1u) : 0u);
				sticky |= V[j] & below;
				// lowest kept bit (cut <= 95 lies in these limbs as well)
				if (cut >= lo && cut < lo + 32) lsb = (V[j] >> (cut - lo)) & 1u;
				const uint32_t keep = cut >= lo + 32 ? 0u : (cut > lo ? ~((1u << (cut - lo)) - 1u) : 0xFFFFFFFFu);
				V[j] &= keep;
			}
			if (rb && (sticky || lsb))
			{
				uint32_t carry = 0u;
#pragma unroll
				for (int j = 0; j < 3; ++j)
				{
					const int lo = 32 * j;
					const uint32_t add = (cut >= lo && cut < lo + 32) ? (1u << (cut - lo)) : 0u;
					const uint64_t t = uint64_t(V[j]) + add + carry;
					V[j] = uint32_t(t);
					carry = uint32_t(t >> 32);
				}
				if (carry)
				{
					// the carry leaves the three lowest limbs (probability ~2^-(96 - cut))
					bool cy = true;
#pragma unroll
					for (int j = 3; j < W; ++j)
					{
						V[j] += cy ? 1u : 0u;
						cy = cy && V[j] == 0u;
					}
				}
			}
			return true;
		}
		// (sgn, V) <- RN(RN((sgn, V) * n) + (csgn, C)); false: outside the window (V is then unusable)
		template <int L>
		QB_HD QB_INLINE bool horner_step(uint32_t* V, uint32_t& sgn, uint32_t n, const uint32_t* C, uint32_t csgn, int prec)
		{
			constexpr int W = L + 2;
			{
				// V *= n: limb j of the result is lo(V[j] n) + hi(V[j-1] n) + carry -- one wide product and one add-with-carry
				// per limb (an accumulating 64-bit product costs two register moves per limb on top)
				uint64_t p = uint64_t(V[0]) * n;
				V[0] = uint32_t(p);
				uint32_t h = uint32_t(p >> 32);
				p = uint64_t(V[1]) * n;
				QB_CY_DECL
				QB_ADD_FIRST(V[1], uint32_t(p), h);
				h = uint32_t(p >> 32);
#pragma unroll
				for (int j = 2; j < W; ++j)
				{
					p = uint64_t(V[j]) * n;
					QB_ADD_NEXT(V[j], uint32_t(p), h);
					h = uint32_t(p >> 32);
				}
				// no carry out of the top limb and h == 0: the frame exponent bounds every intermediate
			}
			if (!round_sig<L>(V, prec)) return false;
			if (((sgn ^ csgn) & SIGN_BIT) == 0u)
			{
				QB_CY_DECL
				QB_ADD_FIRST(V[0], V[0], C[0]);
#pragma unroll
				for (int j = 1; j < W; ++j) QB_ADD_NEXT(V[j], V[j], C[j]);
			}
			else
			{
				uint32_t bw;
				{
					QB_CY_DECL
					QB_SUB_FIRST(V[0], V[0], C[0]);
#pragma unroll
					for (int j = 1; j < W; ++j) QB_SUB_NEXT(V[j], V[j], C[j]);
					QB_BORROW_OUT(bw);
				}
				if (bw)
				{
					// |C| was the larger one: V = -V, the sign is C's
					QB_CY_DECL
					QB_SUB_FIRST(V[0], 0u, V[0]);
#pragma unroll
					for (int j = 1; j < W; ++j) QB_SUB_NEXT(V[j], 0u, V[j]);
					sgn = csgn & SIGN_BIT;
				}
			}
			return round_sig<L>(V, prec);
		}
		// the value as a record; false when V == 0 (the sign of a zero is left to the floating-point path) or outside
		// the window.  V must be rounded (at most prec significant bits).
		template <int L>
		QB_HD QB_INLINE bool to_record(mpf<L>& p, const uint32_t* V, uint32_t sgn, int32_t E)
		{
			constexpr int W = L + 2;
			int q;  // whole limbs of leading zeros
			uint32_t top;
			if (V[W - 1])
			{
				top = V[W - 1];
				q = 0;
			}
			else if (V[W - 2])
			{
				top = V[W - 2];
				q = 1;
			}
			else if (V[W - 3])
			{
				top = V[W - 3];
				q = 2;
			}
			else
				return false;
			const int s = clz32(top);
			// mantissa limb j (of L) = frame limb j + 2 - q and the one below it, shifted left by s
#pragma unroll
			for (int j = 0; j < L; ++j)
			{
				const uint32_t h0 = V[j + 2], h1 = V[j + 1], h2 = V[j], h3 = j >= 1 ? V[j - 1] : 0u;
				const uint32_t hi = q == 0 ? h0 : (q == 1 ? h1 : h2), lo = q == 0 ? h1 : (q == 1 ? h2 : h3);
				p.m[j] = shl_pair(hi, lo, s);
			}
			p.sk = K_REG | (sgn & SIGN_BIT);
			p.exp = E - 32 * q - s;
			return true;
		}
	}  // namespace fx
}  // namespace qb
