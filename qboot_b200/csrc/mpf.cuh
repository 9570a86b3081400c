// qboot_b200 -- fixed-limb multiprecision binary float for sm_100a (and the host side of the same engine).
//
// Replaces the reference's MPFR-backed `mp::real` (include/qboot/mp/real.hpp:87-612) on the block-table hot path.
// Semantics are those the reference relies on: every operation returns the exact result rounded once, to nearest-even,
// at `prec` bits (MPFR_RNDN at mp::global_prec, src/mp/real.cpp:5-6), so a computation that performs the reference's
// operations in the reference's order reproduces its results bit for bit.
//
// One number = sign/kind word + int32 exponent + L little-endian 32-bit limbs, mantissa left aligned (value =
// 0.m * 2^exp with the top bit of m[L-1] set, MPFR's convention); the low 32L - prec bits are zero.
// All routines are thread-serial `__host__ __device__` code over limb arrays addressed through pointers; the L x L
// limb product is the only O(L^2) part and runs as an unrolled product-scanning loop on the IMAD pipe.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define QB_HD __host__ __device__
#define QB_INLINE __forceinline__
// Device code is fully inlined into the kernels: out-of-line device calls made under warp divergence mis-executed on
// sm_100a (ptxas 12.9 passes warp-uniform pointer arguments in uniform registers that a diverged sibling path's callee
// clobbers).  Loops over limbs are kept rolled (QB_ROLL) except the inner product loop, to bound code size.
#define QB_NOINLINE __forceinline__
#define QB_ROLL _Pragma("unroll 1")
#else
#define QB_ROLL
#define QB_HD
#define QB_INLINE inline
#define QB_NOINLINE
#endif

namespace qb
{
	enum : uint32_t
	{
		K_ZERO = 0u,
		K_REG = 1u,
		K_INF = 2u,
		K_NAN = 3u,
		SIGN_BIT = 0x80000000u
	};

	template <int L>
	struct mpf
	{
		uint32_t sk;   // kind | sign << 31
		int32_t exp;   // value = 0.m * 2^exp
		uint32_t m[L];
	};

	// ---- small helpers ---------------------------------------------------------------------------------------------
	QB_HD QB_INLINE int clz32(uint32_t x)
	{
#if defined(__CUDA_ARCH__)
		return __clz(int(x));
#else
		return x ? __builtin_clz(x) : 32;
#endif
	}
	QB_HD QB_INLINE int clz64(uint64_t x)
	{
#if defined(__CUDA_ARCH__)
		return __clzll((long long)x);
#else
		return x ? __builtin_clzll(x) : 64;
#endif
	}
	// (hi:lo) << s, upper word, 0 <= s < 32
	QB_HD QB_INLINE uint32_t shl_pair(uint32_t hi, uint32_t lo, int s)
	{
#if defined(__CUDA_ARCH__)
		return __funnelshift_l(lo, hi, s);
#else
		return s ? (hi << s) | (lo >> (32 - s)) : hi;
#endif
	}
	// (hi:lo) >> s, lower word, 0 <= s < 32
	QB_HD QB_INLINE uint32_t shr_pair(uint32_t hi, uint32_t lo, int s)
	{
#if defined(__CUDA_ARCH__)
		return __funnelshift_r(lo, hi, s);
#else
		return s ? (lo >> s) | (hi << (32 - s)) : lo;
#endif
	}
	template <int L>
	QB_HD QB_INLINE uint32_t kind(const mpf<L>& x)
	{
		return x.sk & 3u;
	}
	template <int L>
	QB_HD QB_INLINE uint32_t sign(const mpf<L>& x)
	{
		return x.sk & SIGN_BIT;
	}
	template <int L>
	QB_HD QB_INLINE bool is_zero(const mpf<L>& x)
	{
		return kind(x) == K_ZERO;
	}
	template <int L>
	QB_HD QB_INLINE bool is_special(const mpf<L>& x)
	{
		return (x.sk & 2u) != 0u;
	}
	template <int L>
	QB_HD QB_INLINE void set_zero(mpf<L>& r, uint32_t sgn = 0u)
	{
		r.sk = K_ZERO | (sgn & SIGN_BIT);
		r.exp = 0;
		for (int i = 0; i < L; ++i) r.m[i] = 0u;
	}
	template <int L>
	QB_HD QB_INLINE void set_nan(mpf<L>& r)
	{
		set_zero(r);
		r.sk = K_NAN;
	}
	template <int L>
	QB_HD QB_INLINE void set_inf(mpf<L>& r, uint32_t sgn)
	{
		set_zero(r);
		r.sk = K_INF | (sgn & SIGN_BIT);
	}
	template <int L>
	QB_HD QB_INLINE void copy(mpf<L>& r, const mpf<L>& a)
	{
		r.sk = a.sk;
		r.exp = a.exp;
		for (int i = 0; i < L; ++i) r.m[i] = a.m[i];
	}
	template <int L>
	QB_HD QB_INLINE void neg(mpf<L>& r)
	{
		if (kind(r) != K_NAN) r.sk ^= SIGN_BIT;
	}

	// ---- rounding --------------------------------------------------------------------------------------------------
	// Round the exact magnitude  X * 2^(e - 32N) + (sticky ? something in (0, 1) units of X's last bit : 0)
	// (X = N little-endian limbs, any normalisation) to `prec` bits, nearest-even, into r.  X == 0 with sticky == 0
	// gives +0 (MPFR: exact cancellation yields +0 under RNDN).
	template <int L, int N>
	QB_HD QB_NOINLINE void round_to(mpf<L>& r, uint32_t sgn, int32_t e, const uint32_t* X, uint32_t sticky, int prec)
	{
		int t = N - 1;
		while (t >= 0 && X[t] == 0u) --t;
		if (t < 0)
		{
			// a pure sticky remainder cannot occur for the callers (they keep the leading bit inside X)
			set_zero(r, 0u);
			return;
		}
		int s = clz32(X[t]);              // bit shift
		int lz = 32 * (N - 1 - t) + s;    // total leading zero bits
		// normalised limb j (0 = least significant of the L kept) comes from X[t - (L-1) + j] and the one below
		int base = t - (L - 1);
		uint32_t below = 0u;  // bits below the kept L limbs, collapsed
		uint32_t guard;       // the limb immediately below the kept ones, after shift
		QB_ROLL
		for (int j = 0; j < L; ++j)
		{
			int i = base + j;
			uint32_t hi = (i >= 0) ? X[i] : 0u;
			uint32_t lo = (i - 1 >= 0) ? X[i - 1] : 0u;
			r.m[j] = shl_pair(hi, lo, s);
		}
		{
			int i = base - 1;
			uint32_t hi = (i >= 0) ? X[i] : 0u;
			uint32_t lo = (i - 1 >= 0) ? X[i - 1] : 0u;
			guard = shl_pair(hi, lo, s);
			// everything below `guard`
			if (i - 1 >= 0 && s && (X[i - 1] << s) != 0u) below = 1u;
			if (i - 1 >= 0 && !s) below |= (X[i - 1] != 0u);  // when s == 0, X[i-1] is entirely below guard
			QB_ROLL
			for (int k = i - 2; k >= 0; --k) below |= (X[k] != 0u);
		}
		uint32_t st = (sticky != 0u) | (below != 0u);
		int drop = 32 * L - prec;  // 0 <= drop < 32 expected (prec > 32 (L-1)), but any drop < 32L works below
		uint32_t rnd;
		if (drop == 0)
		{
			rnd = guard >> 31;
			st |= ((guard << 1) != 0u);
		}
		else
		{
			// drop in [1, 31]: round bit and sticky live in the low bits of m[0]
			st |= (guard != 0u);
			uint32_t lowmask = (1u << drop) - 1u;
			uint32_t low = r.m[0] & lowmask;
			rnd = (low >> (drop - 1)) & 1u;
			st |= ((low & ((1u << (drop - 1)) - 1u)) != 0u);
			r.m[0] &= ~lowmask;
		}
		int32_t ex = e - lz;
		uint32_t ulp = 1u << drop;  // drop < 32
		uint32_t lsb = (r.m[0] >> drop) & 1u;
		if (rnd && (st || lsb))
		{
			uint32_t c = ulp;
			QB_ROLL
			for (int j = 0; j < L; ++j)
			{
				uint32_t v = r.m[j] + c;
				c = (v < c) ? 1u : 0u;
				r.m[j] = v;
			}
			if (c)
			{
				r.m[L - 1] = 0x80000000u;
				ex += 1;
			}
		}
		r.sk = K_REG | (sgn & SIGN_BIT);
		r.exp = ex;
	}

	// ---- limb products ---------------------------------------------------------------------------------------------
	// out[0 .. NA+NB) = a[0..NA) * b[0..NB).  Operand scanning: one rolled loop over the limbs of a, the inner loop over b
	// unrolled with b and a sliding NB-limb accumulator window held in registers (mad.wide.u32 + 64-bit add per limb
	// product); after row i the lowest window limb is final and is stored.  Keeps the code a few hundred instructions.
	template <int NA, int NB>
	QB_HD QB_INLINE void mul_limbs(uint32_t* out, const uint32_t* a, const uint32_t* b)
	{
#if !defined(__CUDA_ARCH__) && defined(__SIZEOF_INT128__) && defined(__BYTE_ORDER__) && __BYTE_ORDER__ == __ORDER_LITTLE_ENDIAN__
		// host (planner, facade arithmetic): the same exact product on 64-bit limbs -- a quarter of the limb products
		if (NA >= 4 && NB >= 4)
		{
			constexpr int A64 = (NA + 1) / 2, B64 = (NB + 1) / 2;
			uint64_t x[A64], y[B64], z[A64 + B64];
			for (int i = 0; i < A64; ++i) x[i] = uint64_t(a[2 * i]) | (2 * i + 1 < NA ? uint64_t(a[2 * i + 1]) << 32 : 0);
			for (int i = 0; i < B64; ++i) y[i] = uint64_t(b[2 * i]) | (2 * i + 1 < NB ? uint64_t(b[2 * i + 1]) << 32 : 0);
			for (int k = 0; k < A64 + B64; ++k) z[k] = 0;
			for (int i = 0; i < A64; ++i)
			{
				uint64_t carry = 0;
				const uint64_t xi = x[i];
				for (int j = 0; j < B64; ++j)
				{
					const unsigned __int128 t = (unsigned __int128)xi * y[j] + z[i + j] + carry;
					z[i + j] = uint64_t(t);
					carry = uint64_t(t >> 64);
				}
				z[i + B64] = carry;
			}
			for (int k = 0; k < NA + NB; ++k) out[k] = uint32_t(z[k >> 1] >> (32 * (k & 1)));
			return;
		}
#endif
		uint32_t rb[NB], acc[NB];
#pragma unroll
		for (int j = 0; j < NB; ++j)
		{
			rb[j] = b[j];
			acc[j] = 0u;
		}
#pragma unroll 1
		for (int i = 0; i < NA; ++i)
		{
			const uint32_t ai = a[i];
			uint64_t t = uint64_t(ai) * rb[0] + acc[0];
			out[i] = uint32_t(t);
			uint32_t carry = uint32_t(t >> 32);
#pragma unroll
			for (int j = 1; j < NB; ++j)
			{
				t = uint64_t(ai) * rb[j] + acc[j] + carry;
				acc[j - 1] = uint32_t(t);
				carry = uint32_t(t >> 32);
			}
			acc[NB - 1] = carry;
		}
#pragma unroll
		for (int j = 0; j < NB; ++j) out[NA + j] = acc[j];
	}

	// r = RN(a * b)
	template <int L>
	QB_HD QB_NOINLINE void mul(mpf<L>& r, const mpf<L>& a, const mpf<L>& b, int prec)
	{
		uint32_t sgn = (a.sk ^ b.sk) & SIGN_BIT;
		if (((a.sk | b.sk) & 2u) != 0u)
		{
			// inf / nan operands: nan unless inf * nonzero
			if (kind(a) == K_NAN || kind(b) == K_NAN || is_zero(a) || is_zero(b))
				set_nan(r);
			else
				set_inf(r, sgn);
			return;
		}
		if (is_zero(a) || is_zero(b))
		{
			set_zero(r, sgn);
			return;
		}
		uint32_t ra[L], rb[L], P[2 * L];
#pragma unroll
		for (int i = 0; i < L; ++i)
		{
			ra[i] = a.m[i];
			rb[i] = b.m[i];
		}
		int32_t e = a.exp + b.exp;
		mul_limbs<L, L>(P, ra, rb);
		round_to<L, 2 * L>(r, sgn, e, P, 0u, prec);
	}

	// ---- exact addition of two normalised N-limb magnitudes with signs ------------------------------------------------
	// out (N+1 limbs, extra limb at the bottom) and the returned sticky describe  sa*A*2^ea + sb*B*2^eb  exactly:
	// value = (-1)^osgn * (OUT + sticky') * 2^(oe - 32(N+1)).  Returns false when the result is exactly zero.
	template <int N>
	QB_HD QB_NOINLINE bool add_core(uint32_t* OUT, uint32_t& osgn, int32_t& oe, uint32_t& ostk, uint32_t sa, int32_t ea,
	                    const uint32_t* A, uint32_t sb, int32_t eb, const uint32_t* B)
	{
		// order by magnitude so that |A*2^ea| >= |B*2^eb|
		bool swap = false;
		if (ea < eb)
			swap = true;
		else if (ea == eb)
		{
			QB_ROLL
			for (int i = N - 1; i >= 0; --i)
				if (A[i] != B[i])
				{
					swap = A[i] < B[i];
					break;
				}
		}
		const uint32_t* X = swap ? B : A;
		const uint32_t* Y = swap ? A : B;
		uint32_t sx = swap ? sb : sa, sy = swap ? sa : sb;
		int32_t ex = swap ? eb : ea, ey = swap ? ea : eb;
		int64_t d64 = int64_t(ex) - int64_t(ey);
		bool subtract = ((sx ^ sy) & SIGN_BIT) != 0u;
		osgn = sx & SIGN_BIT;
		uint32_t stk = 0u;
		// shifted Y into N+1 limbs (index 0 is the extra low limb): position of Y[i] is limb i+1, shifted right by d
		if (d64 >= int64_t(32) * (N + 1))
		{
			// Y lies entirely below the window
			stk = 1u;
			OUT[0] = 0u;
			QB_ROLL
			for (int i = 0; i < N; ++i) OUT[i + 1] = X[i];
			if (subtract)
			{
				// X - tiny: borrow one unit of the extra limb, remainder becomes positive sticky
				uint32_t bw = 1u;
				QB_ROLL
				for (int i = 0; i < N + 1; ++i)
				{
					uint32_t v = OUT[i] - bw;
					bw = (OUT[i] < bw) ? 1u : 0u;
					OUT[i] = v;
				}
			}
			oe = ex;
			ostk = stk;
			return true;
		}
		int d = int(d64);
		int q = d >> 5, s = d & 31;
		// Ysh[k] for k in [0, N]: bits of (0:Y[N-1..0]:0_extra) >> d.  Source limb index in the (N+1)-limb frame:
		// frame[k] = (k >= 1) ? Y[k-1] : 0
		uint32_t carry = 0u;  // carry / borrow
		QB_ROLL
		for (int k = 0; k < N + 1; ++k)
		{
			int lo_i = k + q;      // frame index of low part
			int hi_i = k + q + 1;  // frame index of high part
			uint32_t lo = (lo_i >= 1 && lo_i <= N) ? Y[lo_i - 1] : 0u;
			uint32_t hi = (hi_i >= 1 && hi_i <= N) ? Y[hi_i - 1] : 0u;
			uint32_t y = shr_pair(hi, lo, s);
			uint32_t x = (k >= 1) ? X[k - 1] : 0u;
			if (!subtract)
			{
				uint32_t v = x + y;
				uint32_t c1 = (v < x) ? 1u : 0u;
				uint32_t w = v + carry;
				uint32_t c2 = (w < v) ? 1u : 0u;
				OUT[k] = w;
				carry = c1 | c2;
			}
			else
			{
				uint32_t v = x - y;
				uint32_t b1 = (x < y) ? 1u : 0u;
				uint32_t w = v - carry;
				uint32_t b2 = (v < carry) ? 1u : 0u;
				OUT[k] = w;
				carry = b1 | b2;
			}
		}
		// bits of Y shifted out below the window
		{
			// frame indices [0, q) entirely (frame[0] = 0 contributes nothing) and the low s bits of frame[q]
			QB_ROLL
			for (int i = 1; i < q && i <= N; ++i) stk |= (Y[i - 1] != 0u);
			if (q >= 1 && q <= N && s) stk |= ((Y[q - 1] << (32 - s)) != 0u);
			if (q >= 1 && q <= N && !s) stk |= 0u;  // frame[q] itself is kept when s == 0
		}
		oe = ex;
		if (!subtract)
		{
			if (carry)
			{
				// shift right by one bit, keep the lost bit in sticky
				stk |= (OUT[0] & 1u);
				QB_ROLL
				for (int k = 0; k < N; ++k) OUT[k] = (OUT[k] >> 1) | (OUT[k + 1] << 31);
				OUT[N] = (OUT[N] >> 1) | 0x80000000u;
				oe = ex + 1;
			}
		}
		else
		{
			// |X| >= |Y| so no final borrow.  If bits were shifted out, the true value is OUT - fraction:
			// borrow one unit of the last place and keep a positive sticky.
			if (stk)
			{
				uint32_t bw = 1u;
				QB_ROLL
				for (int k = 0; k < N + 1; ++k)
				{
					uint32_t v = OUT[k] - bw;
					bw = (OUT[k] < bw) ? 1u : 0u;
					OUT[k] = v;
				}
			}
			else
			{
				bool nz = false;
				QB_ROLL
				for (int k = 0; k < N + 1; ++k) nz |= (OUT[k] != 0u);
				if (!nz)
				{
					ostk = 0u;
					return false;
				}
			}
		}
		ostk = stk;
		return true;
	}

	// r = RN(a + b) (negate_b: a - b)
	template <int L>
	QB_HD QB_NOINLINE void add(mpf<L>& r, const mpf<L>& a, const mpf<L>& b, int prec, bool negate_b = false)
	{
		uint32_t sbk = negate_b ? (b.sk ^ SIGN_BIT) : b.sk;
		if (((a.sk | b.sk) & 2u) != 0u)
		{
			if (kind(a) == K_NAN || kind(b) == K_NAN)
				set_nan(r);
			else if (kind(a) == K_INF && kind(b) == K_INF)
			{
				if ((a.sk ^ sbk) & SIGN_BIT)
					set_nan(r);
				else
					set_inf(r, a.sk);
			}
			else if (kind(a) == K_INF)
				set_inf(r, a.sk);
			else
				set_inf(r, sbk);
			return;
		}
		if (is_zero(b))
		{
			if (is_zero(a))
			{
				// (+0) + (-0) = +0 under RNDN; equal signs keep the sign
				set_zero(r, (a.sk & sbk) & SIGN_BIT);
				return;
			}
			if (&r != &a) copy(r, a);
			return;
		}
		if (is_zero(a))
		{
			if (&r != &b) copy(r, b);
			r.sk = sbk;
			return;
		}
		uint32_t OUT[L + 1];
		uint32_t osgn, ostk;
		int32_t oe;
		if (!add_core<L>(OUT, osgn, oe, ostk, a.sk, a.exp, a.m, sbk, b.exp, b.m))
		{
			set_zero(r, 0u);
			return;
		}
		round_to<L, L + 1>(r, osgn, oe, OUT, ostk, prec);
	}
	template <int L>
	QB_HD QB_INLINE void sub(mpf<L>& r, const mpf<L>& a, const mpf<L>& b, int prec)
	{
		add(r, a, b, prec, true);
	}

	// r = RN(a * b + c)   (mpfr_fma)
	template <int L>
	QB_HD QB_NOINLINE void fma(mpf<L>& r, const mpf<L>& a, const mpf<L>& b, const mpf<L>& c, int prec)
	{
		if (((a.sk | b.sk | c.sk) & 2u) != 0u)
		{
			mpf<L> t;
			mul(t, a, b, prec);
			add(r, t, c, prec);
			return;
		}
		uint32_t psgn = (a.sk ^ b.sk) & SIGN_BIT;
		if (is_zero(a) || is_zero(b))
		{
			if (is_zero(c))
				set_zero(r, (psgn & c.sk) & SIGN_BIT);
			else if (&r != &c)
				copy(r, c);
			return;
		}
		if (is_zero(c))
		{
			mul(r, a, b, prec);
			return;
		}
		uint32_t ra[L], rb[L];
		uint32_t P[2 * L], C[2 * L], OUT[2 * L + 1];
#pragma unroll
		for (int i = 0; i < L; ++i)
		{
			ra[i] = a.m[i];
			rb[i] = b.m[i];
		}
		mul_limbs<L, L>(P, ra, rb);
		int32_t pe = a.exp + b.exp;
		if ((P[2 * L - 1] & 0x80000000u) == 0u)
		{
			// normalise (exact: the product has at most 64L - 1 significant bits here)
			QB_ROLL
			for (int i = 2 * L - 1; i > 0; --i) P[i] = (P[i] << 1) | (P[i - 1] >> 31);
			P[0] <<= 1;
			pe -= 1;
		}
		QB_ROLL
		for (int i = 0; i < L; ++i)
		{
			C[i] = 0u;
			C[L + i] = c.m[i];
		}
		uint32_t osgn, ostk;
		int32_t oe;
		if (!add_core<2 * L>(OUT, osgn, oe, ostk, psgn, pe, P, c.sk, c.exp, C))
		{
			set_zero(r, 0u);
			return;
		}
		round_to<L, 2 * L + 1>(r, osgn, oe, OUT, ostk, prec);
	}

	// ---- integer operands ------------------------------------------------------------------------------------------
	// r = exact integer (must fit: |v| < 2^64 <= 2^prec)
	template <int L>
	QB_HD QB_NOINLINE void set_si(mpf<L>& r, int64_t v, int prec)
	{
		if (v == 0)
		{
			set_zero(r, 0u);
			return;
		}
		uint32_t sgn = v < 0 ? SIGN_BIT : 0u;
		uint64_t u = v < 0 ? uint64_t(0) - uint64_t(v) : uint64_t(v);
		uint32_t X[2] = {uint32_t(u), uint32_t(u >> 32)};
		round_to<L, 2>(r, sgn, 64, X, 0u, prec);
	}
	template <int L>
	QB_HD QB_NOINLINE void set_ui_2exp(mpf<L>& r, uint64_t u, int32_t e, int prec)
	{
		if (u == 0)
		{
			set_zero(r, 0u);
			return;
		}
		uint32_t X[2] = {uint32_t(u), uint32_t(u >> 32)};
		round_to<L, 2>(r, 0u, 64 + e, X, 0u, prec);
	}
	// r = RN(a * v)  (mpfr_mul_si / mpfr_mul_ui)
	template <int L>
	QB_HD QB_NOINLINE void mul_si(mpf<L>& r, const mpf<L>& a, int64_t v, int prec)
	{
		uint32_t vs = v < 0 ? SIGN_BIT : 0u;
		uint64_t u = v < 0 ? uint64_t(0) - uint64_t(v) : uint64_t(v);
		uint32_t sgn = (a.sk ^ vs) & SIGN_BIT;
		if (is_special(a))
		{
			if (kind(a) == K_NAN || u == 0)
				set_nan(r);
			else
				set_inf(r, sgn);
			return;
		}
		if (is_zero(a) || u == 0)
		{
			set_zero(r, sgn);
			return;
		}
		uint32_t X[L + 2];
		uint32_t ulo = uint32_t(u), uhi = uint32_t(u >> 32);
		if (uhi == 0u)
		{
			uint64_t c = 0;
			QB_ROLL
			for (int i = 0; i < L; ++i)
			{
				c += uint64_t(a.m[i]) * ulo;
				X[i] = uint32_t(c);
				c >>= 32;
			}
			X[L] = uint32_t(c);
			X[L + 1] = 0u;
		}
		else
		{
			uint32_t b2[2] = {ulo, uhi};
			uint32_t am[L];
			QB_ROLL
			for (int i = 0; i < L; ++i) am[i] = a.m[i];
			mul_limbs<L, 2>(X, am, b2);
		}
		round_to<L, L + 2>(r, sgn, a.exp + 64, X, 0u, prec);
	}
	// r = RN(a + v)
	template <int L>
	QB_HD QB_NOINLINE void add_si(mpf<L>& r, const mpf<L>& a, int64_t v, int prec)
	{
		if (v == 0)
		{
			// MPFR returns the operand unchanged (keeps the sign of a zero)
			if (&r != &a) copy(r, a);
			return;
		}
		mpf<L> t;
		set_si(t, v, prec);
		add(r, a, t, prec);
	}
	// r = RN(v - a)
	template <int L>
	QB_HD QB_NOINLINE void si_sub(mpf<L>& r, int64_t v, const mpf<L>& a, int prec)
	{
		if (v == 0)
		{
			// MPFR: 0 - a is a plain negation
			if (&r != &a) copy(r, a);
			neg(r);
			return;
		}
		if (v < 0)
		{
			// MPFR evaluates v - a for v < 0 as -(a + |v|): an exact cancellation gives -0
			add_si(r, a, -v, prec);
			neg(r);
			return;
		}
		mpf<L> t;
		set_si(t, v, prec);
		add(r, t, a, prec, true);
	}
	// r = a * 2^k (exact)
	template <int L>
	QB_HD QB_INLINE void mul_2si(mpf<L>& r, const mpf<L>& a, int32_t k)
	{
		if (&r != &a) copy(r, a);
		if (kind(r) == K_REG) r.exp += k;
	}

	// ---- short division (divisor < 2^32) ------------------------------------------------------------------------------
	// Q[0..N) = floor(X / d), returns remainder
	template <int N>
	QB_HD QB_INLINE uint32_t divrem_1(uint32_t* Q, const uint32_t* X, uint32_t d)
	{
		uint64_t rem = 0;
		QB_ROLL
		for (int i = N - 1; i >= 0; --i)
		{
			uint64_t cur = (rem << 32) | X[i];
			uint64_t q = cur / d;
			rem = cur - q * d;
			Q[i] = uint32_t(q);
		}
		return uint32_t(rem);
	}
	// r = RN(a / v)  (mpfr_div_si / mpfr_div_ui), |v| < 2^32
	template <int L>
	QB_HD QB_NOINLINE void div_si(mpf<L>& r, const mpf<L>& a, int64_t v, int prec)
	{
		uint32_t vs = v < 0 ? SIGN_BIT : 0u;
		uint64_t u = v < 0 ? uint64_t(0) - uint64_t(v) : uint64_t(v);
		uint32_t sgn = (a.sk ^ vs) & SIGN_BIT;
		if (is_special(a))
		{
			if (kind(a) == K_NAN)
				set_nan(r);
			else
				set_inf(r, sgn);
			return;
		}
		if (u == 0)
		{
			if (is_zero(a))
				set_nan(r);
			else
				set_inf(r, sgn);
			return;
		}
		if (is_zero(a))
		{
			set_zero(r, sgn);
			return;
		}
		// numerator: a.m followed by two zero limbs below  ->  quotient keeps >= 32L + 32 significant bits
		uint32_t X[L + 2], Q[L + 2];
		X[0] = X[1] = 0u;
		QB_ROLL
		for (int i = 0; i < L; ++i) X[i + 2] = a.m[i];
		uint32_t rem = divrem_1<L + 2>(Q, X, uint32_t(u));
		round_to<L, L + 2>(r, sgn, a.exp, Q, rem != 0u, prec);
	}

	// ---- rational operands (num / den, |num| < 2^63, 0 < den < 2^32) --------------------------------------------------
	// r = RN(num / den)  (mpfr_set_q)
	template <int L>
	QB_HD QB_NOINLINE void set_q(mpf<L>& r, int64_t num, uint32_t den, int prec)
	{
		if (num == 0)
		{
			set_zero(r, 0u);
			return;
		}
		uint32_t sgn = num < 0 ? SIGN_BIT : 0u;
		uint64_t u = num < 0 ? uint64_t(0) - uint64_t(num) : uint64_t(num);
		uint32_t X[L + 4], Q[L + 4];
		QB_ROLL
		for (int i = 0; i < L + 2; ++i) X[i] = 0u;
		X[L + 2] = uint32_t(u);
		X[L + 3] = uint32_t(u >> 32);
		uint32_t rem = divrem_1<L + 4>(Q, X, den);
		round_to<L, L + 4>(r, sgn, 64, Q, rem != 0u, prec);
	}
	// r = RN(a * num / den)  (mpfr_mul_q: one rounding)
	template <int L>
	QB_HD QB_NOINLINE void mul_q(mpf<L>& r, const mpf<L>& a, int64_t num, uint32_t den, int prec)
	{
		uint32_t ns = num < 0 ? SIGN_BIT : 0u;
		uint64_t u = num < 0 ? uint64_t(0) - uint64_t(num) : uint64_t(num);
		uint32_t sgn = (a.sk ^ ns) & SIGN_BIT;
		if (is_special(a))
		{
			if (kind(a) == K_NAN || u == 0)
				set_nan(r);
			else
				set_inf(r, sgn);
			return;
		}
		if (is_zero(a) || u == 0)
		{
			set_zero(r, sgn);
			return;
		}
		uint32_t X[L + 4], Q[L + 4];
		uint32_t b2[2] = {uint32_t(u), uint32_t(u >> 32)};
		uint32_t am[L];
		QB_ROLL
		for (int i = 0; i < L; ++i) am[i] = a.m[i];
		X[0] = X[1] = 0u;
		mul_limbs<L, 2>(X + 2, am, b2);
		uint32_t rem = divrem_1<L + 4>(Q, X, den);
		round_to<L, L + 4>(r, sgn, a.exp + 64, Q, rem != 0u, prec);
	}
	// r = RN(a / (num / den)) = RN(a * den / num)  (mpfr_div_q), |num| < 2^32
	template <int L>
	QB_HD QB_NOINLINE void div_q(mpf<L>& r, const mpf<L>& a, int64_t num, uint32_t den, int prec)
	{
		uint32_t ns = num < 0 ? SIGN_BIT : 0u;
		uint64_t u = num < 0 ? uint64_t(0) - uint64_t(num) : uint64_t(num);
		uint32_t sgn = (a.sk ^ ns) & SIGN_BIT;
		if (is_special(a))
		{
			if (kind(a) == K_NAN)
				set_nan(r);
			else
				set_inf(r, sgn);
			return;
		}
		if (u == 0)
		{
			if (is_zero(a))
				set_nan(r);
			else
				set_inf(r, sgn);
			return;
		}
		if (is_zero(a))
		{
			set_zero(r, sgn);
			return;
		}
		uint32_t X[L + 3], Q[L + 3];
		X[0] = X[1] = 0u;
		uint64_t c = 0;
		QB_ROLL
		for (int i = 0; i < L; ++i)
		{
			c += uint64_t(a.m[i]) * den;
			X[i + 2] = uint32_t(c);
			c >>= 32;
		}
		X[L + 2] = uint32_t(c);
		uint32_t rem = divrem_1<L + 3>(Q, X, uint32_t(u));
		round_to<L, L + 3>(r, sgn, a.exp + 32, Q, rem != 0u, prec);
	}
	// r = RN(a + num / den)  (mpfr_add_q: correctly rounded).  Computed as (a * den + num) / den with the numerator
	// exact inside a 2L-limb window; the quotient by den < 2^32 of such a numerator cannot sit within 2^-(32L) of a
	// rounding boundary unless it is exact, so one rounding of the windowed value is the correct rounding.
	template <int L>
	QB_HD QB_NOINLINE void add_q(mpf<L>& r, const mpf<L>& a, int64_t num, uint32_t den, int prec)
	{
		if (is_special(a))
		{
			if (&r != &a) copy(r, a);
			return;
		}
		if (num == 0)
		{
			if (&r != &a) copy(r, a);  // MPFR returns the operand unchanged
			return;
		}
		if (is_zero(a))
		{
			set_q(r, num, den, prec);
			return;
		}
		constexpr int N = 2 * L;
		uint32_t A[N], B[N], OUT[N + 1], Q[N + 3], X[N + 3];
		int32_t ae;
		// A = a.m * den, normalised to N limbs
		{
			uint32_t T[L + 1];
			uint64_t c = 0;
			QB_ROLL
			for (int i = 0; i < L; ++i)
			{
				c += uint64_t(a.m[i]) * den;
				T[i] = uint32_t(c);
				c >>= 32;
			}
			T[L] = uint32_t(c);
			// a.m normalised and den >= 1: top nonzero limb is T[L] or T[L-1]
			int t = T[L] ? L : L - 1;
			int s = clz32(T[t]);
			QB_ROLL
			for (int i = 0; i < N; ++i)
			{
				int src = t - (N - 1) + i;
				uint32_t hi = (src >= 0) ? T[src] : 0u;
				uint32_t lo = (src - 1 >= 0) ? T[src - 1] : 0u;
				A[i] = shl_pair(hi, lo, s);
			}
			// value = 0.A * 2^ae
			// T as integer = a.m * den; a = 0.m * 2^exp  =>  a*den = T * 2^(exp - 32L); T has 32(t+1) - s bits
			// => 0.A * 2^(exp - 32L + 32(t+1) - s)
			ae = a.exp - 32 * L + 32 * (t + 1) - s;
		}
		uint32_t ns = num < 0 ? SIGN_BIT : 0u;
		uint64_t u = num < 0 ? uint64_t(0) - uint64_t(num) : uint64_t(num);
		int nb = 64 - clz64(u);
		uint64_t un = u << (64 - nb);
		QB_ROLL
		for (int i = 0; i < N; ++i) B[i] = 0u;
		B[N - 1] = uint32_t(un >> 32);
		B[N - 2] = uint32_t(un);
		int32_t be = nb;
		uint32_t osgn, ostk;
		int32_t oe;
		if (!add_core<N>(OUT, osgn, oe, ostk, a.sk, ae, A, ns, be, B))
		{
			set_zero(r, 0u);
			return;
		}
		X[0] = X[1] = 0u;
		QB_ROLL
		for (int i = 0; i < N + 1; ++i) X[i + 2] = OUT[i];
		uint32_t rem = divrem_1<N + 3>(Q, X, den);
		round_to<L, N + 3>(r, osgn, oe, Q, (rem != 0u) | ostk, prec);
	}
	template <int L>
	QB_HD QB_INLINE void sub_q(mpf<L>& r, const mpf<L>& a, int64_t num, uint32_t den, int prec)
	{
		add_q(r, a, -num, den, prec);
	}

	// ---- division --------------------------------------------------------------------------------------------------
	// r = RN(a / b): schoolbook long division (Knuth D) producing L+1 quotient limbs and a remainder-nonzero sticky
	template <int L>
	QB_HD QB_NOINLINE void div(mpf<L>& r, const mpf<L>& a, const mpf<L>& b, int prec)
	{
		uint32_t sgn = (a.sk ^ b.sk) & SIGN_BIT;
		if (((a.sk | b.sk) & 2u) != 0u)
		{
			if (kind(a) == K_NAN || kind(b) == K_NAN || (kind(a) == K_INF && kind(b) == K_INF))
				set_nan(r);
			else if (kind(a) == K_INF)
				set_inf(r, sgn);
			else
				set_zero(r, sgn);
			return;
		}
		if (is_zero(b))
		{
			if (is_zero(a))
				set_nan(r);
			else
				set_inf(r, sgn);
			return;
		}
		if (is_zero(a))
		{
			set_zero(r, sgn);
			return;
		}
		// numerator window R[0 .. L] (L+1 limbs, R[L] is the overflow limb); divisor D normalised (top bit set)
		uint32_t R[L + 1], D[L], Q[L + 1];
		QB_ROLL
		for (int i = 0; i < L; ++i) D[i] = b.m[i];
		// ensure numerator < divisor so that every quotient digit fits a limb: if a.m >= b.m use a.m / 2
		bool ge = true;
		QB_ROLL
		for (int i = L - 1; i >= 0; --i)
			if (a.m[i] != b.m[i])
			{
				ge = a.m[i] > b.m[i];
				break;
			}
		int32_t e = a.exp - b.exp;
		uint32_t nextlimb = 0u;  // limb fed in below the window at the first step
		if (ge)
		{
			QB_ROLL
			for (int i = 0; i < L - 1; ++i) R[i] = (a.m[i] >> 1) | (a.m[i + 1] << 31);
			R[L - 1] = a.m[L - 1] >> 1;
			nextlimb = a.m[0] << 31;
			e += 1;
		}
		else
			QB_ROLL
			for (int i = 0; i < L; ++i) R[i] = a.m[i];
		R[L] = 0u;
#if !defined(__CUDA_ARCH__) && defined(__SIZEOF_INT128__) && defined(__BYTE_ORDER__) && __BYTE_ORDER__ == __ORDER_LITTLE_ENDIAN__
		// host (facade arithmetic, planner): the same quotient by Knuth D on 64-bit limbs.  U = (R 2^32 + nextlimb) 2^(32 (L + 1))
		// is divided by V = D (even L) or D 2^32 (odd L: keeps V normalised); for even L that is one 32-bit limb more of
		// quotient than needed, which joins the sticky bit.
		if (L >= 4)
		{
			constexpr int NV = (L + 1) / 2, NU = L + 1, NQ = NU - NV + 1;
			uint64_t V[NV], U[NU + 1], QQ[NQ];
			{
				uint32_t v32[2 * NV], u32[2 * NU];
				for (int i = 0; i < 2 * NV; ++i) v32[i] = 0u;
				for (int i = 0; i < L; ++i) v32[i + (L & 1)] = D[i];
				for (int i = 0; i < 2 * NU; ++i) u32[i] = 0u;
				u32[L + 1] = nextlimb;
				for (int i = 0; i < L; ++i) u32[L + 2 + i] = R[i];
				for (int i = 0; i < NV; ++i) V[i] = uint64_t(v32[2 * i]) | (uint64_t(v32[2 * i + 1]) << 32);
				for (int i = 0; i < NU; ++i) U[i] = uint64_t(u32[2 * i]) | (uint64_t(u32[2 * i + 1]) << 32);
				U[NU] = 0;
			}
			const uint64_t v1 = V[NV - 1], v0 = NV >= 2 ? V[NV - 2] : 0;
			for (int j = NQ - 1; j >= 0; --j)
			{
				// quotient digit from U[j + NV], U[j + NV - 1], U[j + NV - 2] and v1, v0
				const uint64_t u2 = U[j + NV], u1 = U[j + NV - 1], u0 = j + NV >= 2 ? U[j + NV - 2] : 0;
				unsigned __int128 qhat, rhat;
				if (u2 >= v1)
				{
					qhat = ~uint64_t(0);
					rhat = (((unsigned __int128)u2 << 64) | u1) - qhat * v1;
				}
				else
				{
					const unsigned __int128 num = ((unsigned __int128)u2 << 64) | u1;
					qhat = num / v1;
					rhat = num - qhat * v1;
				}
				while ((rhat >> 64) == 0 && qhat * v0 > ((rhat << 64) | u0))
				{
					qhat -= 1;
					rhat += v1;
				}
				uint64_t q = uint64_t(qhat), borrow = 0, carry = 0;
				for (int i = 0; i < NV; ++i)
				{
					const unsigned __int128 p = (unsigned __int128)q * V[i] + carry;
					carry = uint64_t(p >> 64);
					const uint64_t lo = uint64_t(p), cur = U[j + i], d1_ = cur - lo;
					const uint64_t b1 = cur < lo;
					U[j + i] = d1_ - borrow;
					borrow = b1 | (d1_ < borrow);
				}
				{
					const uint64_t cur = U[j + NV], d1_ = cur - carry;
					const uint64_t b1 = cur < carry;
					U[j + NV] = d1_ - borrow;
					borrow = b1 | (d1_ < borrow);
				}
				if (borrow)
				{
					q -= 1;
					uint64_t c = 0;
					for (int i = 0; i < NV; ++i)
					{
						const unsigned __int128 t = (unsigned __int128)U[j + i] + V[i] + c;
						U[j + i] = uint64_t(t);
						c = uint64_t(t >> 64);
					}
					U[j + NV] += c;
				}
				QQ[j] = q;
			}
			uint64_t rem = 0;
			for (int i = 0; i < NV; ++i) rem |= U[i];
			uint32_t st = rem != 0;
			if (L & 1)
				for (int i = 0; i <= L; ++i) Q[i] = uint32_t(QQ[i >> 1] >> (32 * (i & 1)));
			else
			{
				st |= uint32_t(QQ[0]) != 0u;
				for (int i = 0; i <= L; ++i) Q[i] = uint32_t(QQ[(i + 1) >> 1] >> (32 * ((i + 1) & 1)));
			}
			round_to<L, L + 1>(r, sgn, e, Q, st, prec);
			return;
		}
#endif
		const uint32_t d1 = D[L - 1];
		const uint32_t d0 = (L >= 2) ? D[L - 2] : 0u;
		QB_ROLL
		for (int j = L; j >= 0; --j)
		{
			// shift window up by one limb, bringing in the next numerator limb
			QB_ROLL
			for (int i = L; i > 0; --i) R[i] = R[i - 1];
			R[0] = nextlimb;
			nextlimb = 0u;
			// estimate quotient digit from the top two limbs of R and d1 (Knuth step D3)
			uint64_t num2 = (uint64_t(R[L]) << 32) | R[L - 1];
			uint64_t qhat, rhat;
			if (R[L] >= d1)
			{
				qhat = 0xFFFFFFFFull;
				rhat = num2 - qhat * d1;
			}
			else
			{
				qhat = num2 / d1;
				rhat = num2 - qhat * d1;
			}
			uint32_t r2 = (L >= 2) ? R[L - 2] : 0u;
			while (rhat <= 0xFFFFFFFFull && qhat * d0 > ((rhat << 32) | r2))
			{
				qhat -= 1;
				rhat += d1;
			}
			// multiply and subtract
			uint64_t borrow = 0, carry = 0;
			QB_ROLL
			for (int i = 0; i < L; ++i)
			{
				carry += qhat * D[i];
				uint64_t sub = (carry & 0xFFFFFFFFull) + borrow;
				carry >>= 32;
				uint64_t cur = uint64_t(R[i]);
				borrow = (cur < sub) ? 1 : 0;
				R[i] = uint32_t(cur - sub);
			}
			{
				uint64_t sub = carry + borrow;
				uint64_t cur = uint64_t(R[L]);
				borrow = (cur < sub) ? 1 : 0;
				R[L] = uint32_t(cur - sub);
			}
			if (borrow)
			{
				// add back
				qhat -= 1;
				uint64_t c = 0;
				QB_ROLL
				for (int i = 0; i < L; ++i)
				{
					c += uint64_t(R[i]) + D[i];
					R[i] = uint32_t(c);
					c >>= 32;
				}
				R[L] += uint32_t(c);
			}
			Q[j] = uint32_t(qhat);
		}
		uint32_t st = 0u;
		QB_ROLL
		for (int i = 0; i <= L; ++i) st |= R[i];
		round_to<L, L + 1>(r, sgn, e, Q, st != 0u, prec);
	}

	// ---- comparison / predicates -----------------------------------------------------------------------------------
	// sign of a - b for zero/regular operands
	template <int L>
	QB_HD QB_NOINLINE int cmp(const mpf<L>& a, const mpf<L>& b)
	{
		int sa = is_zero(a) ? 0 : (sign(a) ? -1 : 1);
		int sb = is_zero(b) ? 0 : (sign(b) ? -1 : 1);
		if (sa != sb) return sa < sb ? -1 : 1;
		if (sa == 0) return 0;
		int m = 0;
		if (a.exp != b.exp)
			m = a.exp < b.exp ? -1 : 1;
		else
			QB_ROLL
			for (int i = L - 1; i >= 0; --i)
				if (a.m[i] != b.m[i])
				{
					m = a.m[i] < b.m[i] ? -1 : 1;
					break;
				}
		return sa * m;
	}
	// mpfr_integer_p
	template <int L>
	QB_HD QB_NOINLINE bool is_integer(const mpf<L>& a)
	{
		if (is_zero(a)) return true;
		if (kind(a) != K_REG) return false;
		if (a.exp <= 0) return false;
		// bits below position (32L - exp) from the top must be zero
		int64_t frac_bits = int64_t(32) * L - a.exp;  // number of low bits that are fractional
		if (frac_bits <= 0) return true;
		int q = int(frac_bits >> 5), s = int(frac_bits & 31);
		QB_ROLL
		for (int i = 0; i < q && i < L; ++i)
			if (a.m[i]) return false;
		if (q < L && s && (a.m[q] & ((1u << s) - 1u))) return false;
		return true;
	}
}  // namespace qb
