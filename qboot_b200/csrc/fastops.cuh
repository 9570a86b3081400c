// qboot_b200 -- register-resident fast paths of the hot arithmetic (mul, fma, add, mul by a 32-bit integer).
//
// Same contract as mpf.cuh (exact result, rounded once to nearest-even at `prec` bits) and bit-identical results; the
// difference is where the limbs live.  The generic routines of mpf.cuh build the exact 2L-limb intermediate in
// thread-local arrays and index them dynamically, which on the GPU means stack (local-memory) traffic: the first ncu
// capture showed the first derivative kernel moving 22 GB through DRAM for 0.2 GB of algorithmic data.  Here every intermediate is a
// statically indexed array that the compiler keeps in registers:
//
//   * the L x L product keeps only its top L + 3 limbs and an OR of everything below (`mul_top`): a rolled loop over the
//     rows, the multiplier and a sliding accumulator window in registers, IMAD.WIDE + carry add per limb product;
//   * alignment for the addition is a logarithmic limb shifter on registers (predicated selects) plus a funnel shift;
//     any exponent distance is handled (every other series coefficient is a rounding residue ~2^-prec smaller than
//     its neighbours, so the Horner chains see distances of ~prec bits on every second step).  Only deep cancellations
//     (>= 63 leading bits) and a few undecidable remainder patterns fall back to the generic routine (same result by
//     construction);
//   * rounding works on the L + 2 limb frame directly.
#pragma once

#include "mpf.cuh"

namespace qb
{
	namespace fast
	{
		// ---- cold paths ---------------------------------------------------------------------------------------------------
		// The generic routines of mpf.cuh index their operands dynamically, so any record handed to them by reference is
		// forced into thread-local MEMORY for its whole lifetime -- also on the hot path, where the call is never taken
		// (measured: 96 local loads / stores per mul_small, 130 + 154 per fma, 13 % of all instructions of k_series).  The
		// fall-backs therefore work on copies made inside the cold branch; the caller's records stay in registers.
		// QB_COLD_COPY = 0 hands the caller's records over directly.  Which is faster is decided per kernel group by measurement
		// (kernel_inst.cu): the copies win in k_series (170 registers: -19 %) and k_scale (-2 %), and lose in k_horner (+6 %),
		// the transverse recursion (+11 %) and the coefficients (+21 %), whose operands are better off in L1-resident local
		// memory than competing for registers (profiles/r2_summary.md).
#ifndef QB_COLD_COPY
#define QB_COLD_COPY 0
#endif
		template <int L>
		QB_HD QB_INLINE void cold_mul(mpf<L>& r, const mpf<L>& a, const mpf<L>& b, int prec)
		{
#if !QB_COLD_COPY
			qb::mul(r, a, b, prec);
			return;
#endif
			mpf<L> a2, b2, r2;
			copy(a2, a);
			copy(b2, b);
			qb::mul(r2, a2, b2, prec);
			copy(r, r2);
		}
		template <int L>
		QB_HD QB_INLINE void cold_fma(mpf<L>& r, const mpf<L>& a, const mpf<L>& b, const mpf<L>& c, int prec)
		{
#if !QB_COLD_COPY
			qb::fma(r, a, b, c, prec);
			return;
#endif
			mpf<L> a2, b2, c2, r2;
			copy(a2, a);
			copy(b2, b);
			copy(c2, c);
			qb::fma(r2, a2, b2, c2, prec);
			copy(r, r2);
		}
		template <int L>
		QB_HD QB_INLINE void cold_add(mpf<L>& r, const mpf<L>& a, const mpf<L>& b, int prec, bool negate_b)
		{
#if !QB_COLD_COPY
			qb::add(r, a, b, prec, negate_b);
			return;
#endif
			mpf<L> a2, b2, r2;
			copy(a2, a);
			copy(b2, b);
			qb::add(r2, a2, b2, prec, negate_b);
			copy(r, r2);
		}
		template <int L>
		QB_HD QB_INLINE void cold_div(mpf<L>& r, const mpf<L>& a, const mpf<L>& b, int prec)
		{
#if !QB_COLD_COPY
			qb::div(r, a, b, prec);
			return;
#endif
			mpf<L> a2, b2, r2;
			copy(a2, a);
			copy(b2, b);
			qb::div(r2, a2, b2, prec);
			copy(r, r2);
		}
		template <int L>
		QB_HD QB_INLINE void cold_mul_si(mpf<L>& r, const mpf<L>& a, int64_t v, int prec)
		{
#if !QB_COLD_COPY
			qb::mul_si(r, a, v, prec);
			return;
#endif
			mpf<L> a2, r2;
			copy(a2, a);
			qb::mul_si(r2, a2, v, prec);
			copy(r, r2);
		}
		template <int L>
		QB_HD QB_INLINE void cold_add_si(mpf<L>& r, const mpf<L>& a, int64_t v, int prec)
		{
#if !QB_COLD_COPY
			qb::add_si(r, a, v, prec);
			return;
#endif
			mpf<L> a2, r2;
			copy(a2, a);
			qb::add_si(r2, a2, v, prec);
			copy(r, r2);
		}

		// ---- carry-chain primitives ------------------------------------------------------------------------------------
		// On the device these are PTX mad.lo.cc / madc.hi.cc pairs, which ptxas fuses into ONE IMAD.WIDE.U32.X per limb
		// product with the carry travelling in a predicate; the host build (and -DQB_EMULATE_CARRY_CHAINS, used by the
		// tests to validate the dataflow) emulates the carry flag in a variable.
#if defined(__CUDA_ARCH__) && !defined(QB_EMULATE_CARRY_CHAINS)
#define QB_CY_DECL
#define QB_MAD_FIRST(lo, hi, x, y) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(x), "r"(y))
#define QB_MAD_NEXT(lo, hi, x, y) asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(x), "r"(y))
#define QB_ADDC_CC(v) asm volatile("addc.cc.u32 %0, %0, 0;" : "+r"(v))
#define QB_ADDC_END(v) asm volatile("addc.u32 %0, %0, 0;" : "+r"(v))
#define QB_ADD_FIRST(d, x, y) asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(x), "r"(y))
#define QB_ADD_NEXT(d, x, y) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(x), "r"(y))
#define QB_SUB_FIRST(d, x, y) asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(x), "r"(y))
#define QB_SUB_NEXT(d, x, y) asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(x), "r"(y))
#define QB_SET_BORROW(bw)                                                     \
	{                                                                         \
		uint32_t dummy_;                                                      \
		asm volatile("sub.cc.u32 %0, 0, %1;" : "=r"(dummy_) : "r"(bw));       \
	}
#define QB_SET_CARRY(cin)                                                              \
	{                                                                                  \
		uint32_t dummy_;                                                               \
		asm volatile("add.cc.u32 %0, %1, 0xffffffff;" : "=r"(dummy_) : "r"(cin));      \
	}
#define QB_CARRY_OUT(v) asm volatile("addc.u32 %0, 0, 0;" : "=r"(v))
// after a subtraction chain: v = 0 - 0 - borrow  (nonzero iff the chain borrowed)
#define QB_BORROW_OUT(v) asm volatile("subc.u32 %0, 0, 0;" : "=r"(v))
#else
#define QB_CY_DECL uint32_t cy_ = 0u;
#define QB_MAD_FIRST(lo, hi, x, y)                                                      \
	{                                                                                   \
		uint64_t t_ = uint64_t(x) * uint64_t(y);                                        \
		uint64_t l_ = uint64_t(lo) + uint32_t(t_);                                      \
		uint64_t h_ = uint64_t(hi) + uint32_t(t_ >> 32) + uint32_t(l_ >> 32);           \
		lo = uint32_t(l_);                                                              \
		hi = uint32_t(h_);                                                              \
		cy_ = uint32_t(h_ >> 32);                                                       \
	}
#define QB_MAD_NEXT(lo, hi, x, y)                                                       \
	{                                                                                   \
		uint64_t t_ = uint64_t(x) * uint64_t(y);                                        \
		uint64_t l_ = uint64_t(lo) + uint32_t(t_) + cy_;                                \
		uint64_t h_ = uint64_t(hi) + uint32_t(t_ >> 32) + uint32_t(l_ >> 32);           \
		lo = uint32_t(l_);                                                              \
		hi = uint32_t(h_);                                                              \
		cy_ = uint32_t(h_ >> 32);                                                       \
	}
#define QB_ADDC_CC(v)                        \
	{                                        \
		uint64_t s_ = uint64_t(v) + cy_;     \
		v = uint32_t(s_);                    \
		cy_ = uint32_t(s_ >> 32);            \
	}
#define QB_ADDC_END(v) \
	{                  \
		v += cy_;      \
		cy_ = 0u;      \
	}
#define QB_ADD_FIRST(d, x, y)                        \
	{                                                \
		uint64_t s_ = uint64_t(x) + uint64_t(y);     \
		d = uint32_t(s_);                            \
		cy_ = uint32_t(s_ >> 32);                    \
	}
#define QB_ADD_NEXT(d, x, y)                               \
	{                                                      \
		uint64_t s_ = uint64_t(x) + uint64_t(y) + cy_;     \
		d = uint32_t(s_);                                  \
		cy_ = uint32_t(s_ >> 32);                          \
	}
// subtraction chains: cy_ holds the borrow
#define QB_SUB_FIRST(d, x, y)                        \
	{                                                \
		uint64_t s_ = uint64_t(x) - uint64_t(y);     \
		d = uint32_t(s_);                            \
		cy_ = uint32_t(s_ >> 63);                    \
	}
#define QB_SUB_NEXT(d, x, y)                               \
	{                                                      \
		uint64_t s_ = uint64_t(x) - uint64_t(y) - cy_;     \
		d = uint32_t(s_);                                  \
		cy_ = uint32_t(s_ >> 63);                          \
	}
#define QB_SET_BORROW(bw) cy_ = (bw) ? 1u : 0u;
#define QB_SET_CARRY(cin) cy_ = (cin) ? 1u : 0u;
#define QB_CARRY_OUT(v) v = cy_;
#define QB_BORROW_OUT(v) v = cy_ ? 0xFFFFFFFFu : 0u;
#endif

		// X[0 .. L+3) = top L+3 limbs of a * b, st = OR of the L-3 limbs below.  a may live anywhere (two limbs are read
		// per iteration), rb must be registers.
		//
		// Two rows of the schoolbook product per iteration of a rolled loop.  Limb products whose position is even
		// relative to the window go to 64-bit aligned pairs of EV, the others to pairs of OD, so that every product is a
		// single multiply-add-with-carry into an aligned register pair (one IMAD.WIDE.U32.X).  After the two rows OD is
		// folded into EV with one add-with-carry chain that also slides the window up by two limbs:
		//   EV[m] <-> position i + m,  OD[m] <-> position i + 1 + m
		//   row i   : a_i b_k (k even) -> EV[k], EV[k+1];   a_i b_k (k odd) -> OD[k-1], OD[k]
		//   row i+1 : a_{i+1} b_k (k even) -> OD[k], OD[k+1];   (k odd) -> EV[k+1], EV[k+2]
		// AS: distance in words between consecutive limbs of a (1 for a record; the block size when the kernel stages a
		// in shared memory, limb-major, so that the per-row fetch never touches thread-local memory).
		template <int L, int AS = 1>
		QB_HD QB_INLINE void mul_top(uint32_t* X, uint32_t& st, const uint32_t* a, const uint32_t* rb)
		{
			constexpr int LP = (L + 1) & ~1;  // even
			uint32_t b[LP], EV[LP + 4], OD[LP + 2];
#pragma unroll
			for (int j = 0; j < LP; ++j) b[j] = j < L ? rb[j] : 0u;
#pragma unroll
			for (int j = 0; j < LP + 4; ++j) EV[j] = 0u;
			uint32_t g0 = 0u, g1 = 0u, g2 = 0u, g3 = 0u, low = 0u;
			// the two multiplier limbs of the NEXT row pair are fetched one iteration ahead: a may live in global or local
			// memory and the rolled loop would otherwise expose that load's latency once per row pair
			uint32_t n0 = a[0], n1 = (1 < L) ? a[AS] : 0u;
#pragma unroll 1
			for (int i = 0; i < LP; i += 2)
			{
				const uint32_t a0 = n0, a1 = n1;
				if (i + 2 < LP)
				{
					n0 = a[(i + 2) * AS];
					n1 = (i + 3 < L) ? a[(i + 3) * AS] : 0u;
				}
				QB_CY_DECL
				// row i, even columns -> EV
				QB_MAD_FIRST(EV[0], EV[1], a0, b[0]);
#pragma unroll
				for (int k = 2; k < LP; k += 2) QB_MAD_NEXT(EV[k], EV[k + 1], a0, b[k]);
				QB_ADDC_CC(EV[LP]);
				QB_ADDC_END(EV[LP + 1]);
				// row i, odd columns -> OD (fresh)
#pragma unroll
				for (int k = 0; k < LP; k += 2)
				{
					uint64_t t = uint64_t(a0) * uint64_t(b[k + 1]);
					OD[k] = uint32_t(t);
					OD[k + 1] = uint32_t(t >> 32);
				}
				OD[LP] = 0u;
				OD[LP + 1] = 0u;
				// row i+1, even columns -> OD
				QB_MAD_FIRST(OD[0], OD[1], a1, b[0]);
#pragma unroll
				for (int k = 2; k < LP; k += 2) QB_MAD_NEXT(OD[k], OD[k + 1], a1, b[k]);
				QB_ADDC_END(OD[LP]);
				// row i+1, odd columns -> EV shifted by two
				QB_MAD_FIRST(EV[2], EV[3], a1, b[1]);
#pragma unroll
				for (int k = 2; k < LP; k += 2) QB_MAD_NEXT(EV[k + 2], EV[k + 3], a1, b[k + 1]);
				QB_ADDC_CC(EV[LP + 2]);
				QB_ADDC_END(EV[LP + 3]);
				// retire positions i and i+1, fold OD into EV and slide the window by two limbs
				low |= g0 | g1;
				g0 = g2;
				g1 = g3;
				g2 = EV[0];
				QB_ADD_FIRST(g3, EV[1], OD[0]);
#pragma unroll
				for (int m = 0; m < LP + 2; ++m)
				{
					const uint32_t o = (m + 1 < LP + 2) ? OD[m + 1] : 0u;
					QB_ADD_NEXT(EV[m], EV[m + 2], o);
				}
				EV[LP + 2] = 0u;
				EV[LP + 3] = 0u;
			}
			if (L % 2 == 0)
			{
				// emitted positions L-4 .. L-1 are g0 .. g3; EV[j] <-> position L + j
				low |= g0;
				X[0] = g1;
				X[1] = g2;
				X[2] = g3;
#pragma unroll
				for (int j = 0; j < L; ++j) X[3 + j] = EV[j];
			}
			else
			{
				// LP = L + 1: emitted positions L-3 .. L are g0 .. g3; EV[j] <-> position L + 1 + j
				X[0] = g0;
				X[1] = g1;
				X[2] = g2;
				X[3] = g3;
#pragma unroll
				for (int j = 0; j + 4 < L + 3; ++j) X[4 + j] = EV[j];
			}
			st = low;
		}

		// Short product: the same frame as mul_top from fewer limb products.  Products a_i b_k whose position i + k lies
		// below T = L - 5 cannot reach the frame except through carries, so most of them are skipped: the rows are cut into
		// segments of R = 8, and the rows of a segment only visit the columns k >= kmin(segment) (a static staircase that
		// covers every product at position >= T).  With S the exact sum of the visited products, the frame and the two
		// limbs f4, f5 below it (positions L-4, L-5) are read off S, and
		//     a b  =  [frame | f4 | f5]  +  r,     0 <= r < L units of f4
		// (skipped products < (L-5)(1+2^-31) units of f4, limbs of S dropped below f5 < one unit of f5).  Hence the frame
		// is exact unless f4 >= 2^32 - L, and the remainder below the frame is known to be nonzero when f4 | f5 != 0.  In
		// the two undecidable cases (probability ~L 2^-32 for full-width operands) the function returns false and the caller
		// takes the full product.  A short multiplicand b (all limbs below position L - 5 zero) is recognised: then nothing
		// nonzero was skipped and f4 = f5 = 0 means an exact frame.  st: as mul_top.
#ifndef QB_SHORT_ROWS
#define QB_SHORT_ROWS 8  // rows per staircase segment (even)
#endif
		template <int L>
		struct ShortPlan
		{
			static constexpr int LP = (L + 1) & ~1, T = L - 5, R = QB_SHORT_ROWS, NSEG = (LP + R - 1) / R;
			static constexpr int i0(int s) { return s * R; }
			static constexpr int i1(int s) { return (s * R + R < LP) ? s * R + R : LP; }
			static constexpr int kmin(int s) { return T - (i1(s) - 1) <= 0 ? 0 : ((T - (i1(s) - 1)) & ~1); }
			// lowest product position visited by segments 0 .. s
			static constexpr int abs_low(int s) { return s == 0 ? i0(0) + kmin(0) : (i0(s) + kmin(s) < abs_low(s - 1) ? i0(s) + kmin(s) : abs_low(s - 1)); }
			static constexpr bool retire(int s) { return (i1(s) - 1) >= (T - 1); }
			// window entries below jlo are zero throughout segment s
			static constexpr int jlo(int s) { return (retire(s) || abs_low(s) - (i1(s) - 2) <= 0) ? 0 : ((abs_low(s) - (i1(s) - 2)) & ~1); }
		};

		// rows [i0, i1) of segment S, then the following segments
		template <int L, int S, int AS>
		QB_HD QB_INLINE void mul_short_seg(uint32_t (&EV)[((L + 1) & ~1) + 4], uint32_t (&OD)[((L + 1) & ~1) + 2], const uint32_t (&b)[(L + 1) & ~1],
		                                   const uint32_t* a, uint32_t& n0, uint32_t& n1, uint32_t (&g)[6])
		{
			using PL = ShortPlan<L>;
			constexpr int LP = PL::LP, i0 = PL::i0(S), i1 = PL::i1(S), kmin = PL::kmin(S), jlo = PL::jlo(S);
			constexpr bool retire = PL::retire(S);
			constexpr int ms = jlo >= 2 ? jlo - 2 : 0;  // first window entry the fold has to produce
			static_assert(kmin % 2 == 0 && jlo % 2 == 0 && (!retire || kmin == 0) && jlo <= kmin, "short-product staircase");
#pragma unroll 1
			for (int i = i0; i < i1; i += 2)
			{
				const uint32_t a0 = n0, a1 = n1;
				if (i + 2 < LP)
				{
					n0 = a[(i + 2) * AS];
					n1 = (i + 3 < L) ? a[(i + 3) * AS] : 0u;
				}
				QB_CY_DECL
				// row i, even columns -> EV
				QB_MAD_FIRST(EV[kmin], EV[kmin + 1], a0, b[kmin]);
#pragma unroll
				for (int k = kmin + 2; k < LP; k += 2) QB_MAD_NEXT(EV[k], EV[k + 1], a0, b[k]);
				QB_ADDC_CC(EV[LP]);
				QB_ADDC_END(EV[LP + 1]);
				// row i, odd columns -> OD (fresh)
#pragma unroll
				for (int k = kmin; k < LP; k += 2)
				{
					uint64_t t = uint64_t(a0) * uint64_t(b[k + 1]);
					OD[k] = uint32_t(t);
					OD[k + 1] = uint32_t(t >> 32);
				}
				OD[LP] = 0u;
				OD[LP + 1] = 0u;
				// row i+1, even columns -> OD
				QB_MAD_FIRST(OD[kmin], OD[kmin + 1], a1, b[kmin]);
#pragma unroll
				for (int k = kmin + 2; k < LP; k += 2) QB_MAD_NEXT(OD[k], OD[k + 1], a1, b[k]);
				QB_ADDC_END(OD[LP]);
				// row i+1, odd columns -> EV shifted by two
				QB_MAD_FIRST(EV[kmin + 2], EV[kmin + 3], a1, b[kmin + 1]);
#pragma unroll
				for (int k = kmin + 2; k < LP; k += 2) QB_MAD_NEXT(EV[k + 2], EV[k + 3], a1, b[k + 1]);
				QB_ADDC_CC(EV[LP + 2]);
				QB_ADDC_END(EV[LP + 3]);
				// fold OD into EV and slide the window by two limbs; positions i and i+1 leave the window
				if (retire)
				{
					g[4] = g[0];
					g[5] = g[1];
					g[0] = g[2];
					g[1] = g[3];
					g[2] = EV[0];
					QB_ADD_FIRST(g[3], EV[1], OD[0]);
#pragma unroll
					for (int m = 0; m < LP + 2; ++m)
					{
						const uint32_t o = (m + 1 < LP + 2) ? OD[m + 1] : 0u;
						QB_ADD_NEXT(EV[m], EV[m + 2], o);
					}
				}
				else if (ms == 0)
				{
					// positions i, i+1 are below everything that is read later, but their carry is kept
					uint32_t dump;
					QB_ADD_FIRST(dump, EV[1], (0 >= kmin ? OD[0] : 0u));
					(void)dump;
#pragma unroll
					for (int m = 0; m < LP + 2; ++m)
					{
						const uint32_t o = (m + 1 >= kmin && m + 1 < LP + 2) ? OD[m + 1] : 0u;
						QB_ADD_NEXT(EV[m], EV[m + 2], o);
					}
				}
				else
				{
					// entries below jlo are zero (and OD below kmin >= jlo is not written in this segment)
					QB_ADD_FIRST(EV[ms], EV[ms + 2], (ms + 1 >= kmin ? OD[ms + 1] : 0u));
#pragma unroll
					for (int m = ms + 1; m < LP + 2; ++m)
					{
						const uint32_t o = (m + 1 >= kmin && m + 1 < LP + 2) ? OD[m + 1] : 0u;
						QB_ADD_NEXT(EV[m], EV[m + 2], o);
					}
				}
				EV[LP + 2] = 0u;
				EV[LP + 3] = 0u;
			}
			if constexpr (S + 1 < PL::NSEG) mul_short_seg<L, S + 1, AS>(EV, OD, b, a, n0, n1, g);
		}

		template <int L, int AS = 1>
		QB_HD QB_INLINE bool mul_top_short(uint32_t* X, uint32_t& st, const uint32_t* a, const uint32_t* rb)
		{
			constexpr int LP = (L + 1) & ~1;  // even
			uint32_t b[LP], EV[LP + 4], OD[LP + 2];
#pragma unroll
			for (int j = 0; j < LP; ++j) b[j] = j < L ? rb[j] : 0u;
#pragma unroll
			for (int j = 0; j < LP + 4; ++j) EV[j] = 0u;
#pragma unroll
			for (int j = 0; j < LP + 2; ++j) OD[j] = 0u;
			uint32_t g[6] = {0u, 0u, 0u, 0u, 0u, 0u};  // g[0..3]: the four positions below the window; g[4], g[5]: the two below those
			uint32_t n0 = a[0], n1 = (1 < L) ? a[AS] : 0u;
			mul_short_seg<L, 0, AS>(EV, OD, b, a, n0, n1, g);
			uint32_t f4, f5;
			if (L % 2 == 0)
			{
				// positions L-4 .. L-1 are g[0..3], L-6 and L-5 are g[4], g[5]; EV[j] <-> position L + j
				f4 = g[0];
				f5 = g[5];
				X[0] = g[1];
				X[1] = g[2];
				X[2] = g[3];
#pragma unroll
				for (int j = 0; j < L; ++j) X[3 + j] = EV[j];
			}
			else
			{
				// LP = L + 1: positions L-3 .. L are g[0..3], L-5 and L-4 are g[4], g[5]; EV[j] <-> position L + 1 + j
				f4 = g[5];
				f5 = g[4];
				X[0] = g[0];
				X[1] = g[1];
				X[2] = g[2];
				X[3] = g[3];
#pragma unroll
				for (int j = 0; j + 4 < L + 3; ++j) X[4 + j] = EV[j];
			}
			st = f4 | f5;
			if (st == 0u)
			{
				// Nothing is known about the remainder -- unless b is short (a machine number, a decimal with few digits, ...):
				// when its limbs below position L - 5 are all zero every skipped product vanishes and the product is exact
				// from position L - 5 up, so the remainder below the frame really is zero.
				uint32_t lowb = 0u;
#pragma unroll
				for (int j = 0; j + 5 < L; ++j) lowb |= b[j];
				return lowb == 0u;
			}
			return f4 < 0xFFFFFFFFu - uint32_t(L);
		}

		// Round the frame X[0 .. L+2) (value 0.X * 2^e, plus a positive remainder below the last bit when st != 0) into r.
		// Requires fewer than 64 leading zero bits with X[L+1] | X[L] != 0; returns false otherwise (nothing written).
		template <int L>
		QB_HD QB_INLINE bool round_frame(mpf<L>& r, uint32_t sgn, int32_t e, const uint32_t* X, uint32_t st, int prec)
		{
			constexpr int N = L + 2;
			const uint32_t top = X[N - 1], nxt = X[N - 2];
			if ((top | nxt) == 0u) return false;
			const bool lq = top == 0u;
			const int lr = clz32(lq ? nxt : top);
			const int lz = (lq ? 32 : 0) + lr;
			// normalised limb k comes from X[k - lq] and the limb below it
			uint32_t Y[N];
			if (lz == 0)
			{
#pragma unroll
				for (int k = 0; k < N; ++k) Y[k] = X[k];
			}
			else
			{
#pragma unroll
				for (int k = 0; k < N; ++k)
				{
					uint32_t h0 = X[k], h1 = (k >= 1) ? X[k - 1] : 0u, h2 = (k >= 2) ? X[k - 2] : 0u;
					uint32_t hi = lq ? h1 : h0, lo = lq ? h2 : h1;
					Y[k] = shl_pair(hi, lo, lr);
				}
			}
			uint32_t guard = Y[1];
			uint32_t sticky = (st != 0u) | (Y[0] != 0u);
			const int drop = 32 * L - prec;
			uint32_t rnd;
			uint32_t m0 = Y[2];
			if (drop == 0)
			{
				rnd = guard >> 31;
				sticky |= ((guard << 1) != 0u);
			}
			else
			{
				sticky |= (guard != 0u);
				uint32_t lowmask = (1u << drop) - 1u;
				uint32_t low = m0 & lowmask;
				rnd = (low >> (drop - 1)) & 1u;
				sticky |= ((low & ((1u << (drop - 1)) - 1u)) != 0u);
				m0 &= ~lowmask;
			}
			uint32_t lsb = (m0 >> drop) & 1u;
			const uint32_t inc = (rnd && (sticky || lsb)) ? (1u << drop) : 0u;
			int32_t ex = e - lz;
			uint32_t cout;
			{
				QB_CY_DECL
				QB_ADD_FIRST(r.m[0], m0, inc);
#pragma unroll
				for (int j = 1; j < L; ++j) QB_ADD_NEXT(r.m[j], Y[j + 2], 0u);
				QB_CARRY_OUT(cout);
			}
			if (cout)
			{
				r.m[L - 1] = 0x80000000u;
				ex += 1;
			}
			r.sk = K_REG | (sgn & SIGN_BIT);
			r.exp = ex;
			return true;
		}

		// V (N limbs, registers) >>= 32 q + s; `lost` collects every bit shifted out.  Logarithmic limb shifter on registers
		// (static indices, predicated selects) followed by a funnel shift.
		template <int N>
		QB_HD QB_INLINE void shr_frame(uint32_t* V, int q, int s, uint32_t& lost)
		{
			if (q >= N)
			{
#pragma unroll
				for (int k = 0; k < N; ++k)
				{
					lost |= V[k];
					V[k] = 0u;
				}
				return;
			}
			if (q <= 2)
			{
				// common case (operands of comparable size): static three-way select
				lost |= (q >= 1 ? V[0] : 0u) | (q >= 2 ? V[1] : 0u);
#pragma unroll
				for (int k = 0; k < N; ++k)
				{
					const uint32_t v1 = (k + 1 < N) ? V[k + 1] : 0u, v2 = (k + 2 < N) ? V[k + 2] : 0u;
					V[k] = q == 0 ? V[k] : (q == 1 ? v1 : v2);
				}
			}
			else
#pragma unroll
			for (int b = 1; b < N; b <<= 1)
			{
				const bool on = (q & b) != 0;
#pragma unroll
				for (int k = 0; k < N; ++k)
				{
					if (k < b) lost |= on ? V[k] : 0u;
					const uint32_t up = (k + b < N) ? V[k + b] : 0u;
					V[k] = on ? up : V[k];
				}
			}
			if (s)
			{
				lost |= V[0] << (32 - s);
#pragma unroll
				for (int k = 0; k < N; ++k)
				{
					const uint32_t hi = (k + 1 < N) ? V[k + 1] : 0u;
					V[k] = shr_pair(hi, V[k], s);
				}
			}
		}

		// r = RN( (-1)^sx 0.X 2^ex (+ st)  +  c ),  X normalised (top bit set), N = L + 2 limbs, c regular and nonzero.
		// Returns false in the rare cases the frame cannot decide the rounding (deep cancellation, or remainders of
		// both operands with an undetermined net sign next to an all-zero / all-one guard limb).
		template <int L>
		QB_HD QB_INLINE bool addround(mpf<L>& r, uint32_t sx, int32_t ex, const uint32_t* X, uint32_t stX, const mpf<L>& c,
		                              int prec)
		{
			constexpr int N = L + 2;
			const int64_t d64 = int64_t(ex) - int64_t(c.exp);
			uint32_t B[N], Sm[N];              // bigger / smaller magnitude, aligned in one frame
			uint32_t stB, stS;                 // remainders below the frame, each extending its own magnitude
			uint32_t sb, ssm;
			int32_t eb;
			if (d64 >= 0)
			{
				// X keeps its place, c moves right by d
#pragma unroll
				for (int k = 0; k < N; ++k)
				{
					B[k] = X[k];
					Sm[k] = (k >= 2) ? c.m[k - 2] : 0u;
				}
				stB = stX != 0u;
				uint32_t lost = 0u;
				const int64_t dd = d64 > int64_t(32) * N ? int64_t(32) * N : d64;
				shr_frame<N>(Sm, int(dd >> 5), int(dd & 31), lost);
				stS = lost != 0u;
				sb = sx;
				ssm = c.sk;
				eb = ex;
				if (d64 == 0)
				{
					// equal exponents: order by magnitude (both top bits set)
					bool swap = false, decided = false;
#pragma unroll
					for (int k = N - 1; k >= 0; --k)
						if (!decided && B[k] != Sm[k])
						{
							swap = B[k] < Sm[k];
							decided = true;
						}
					if (swap)
					{
#pragma unroll
						for (int k = 0; k < N; ++k)
						{
							uint32_t t = B[k];
							B[k] = Sm[k];
							Sm[k] = t;
						}
						sb = c.sk;
						ssm = sx;
						uint32_t t = stB;
						stB = stS;
						stS = t;
					}
				}
			}
			else
			{
				// c keeps the top of the frame, X moves right by -d
#pragma unroll
				for (int k = 0; k < N; ++k)
				{
					Sm[k] = X[k];
					B[k] = (k >= 2) ? c.m[k - 2] : 0u;
				}
				uint32_t lost = stX;
				const int64_t dd = -d64 > int64_t(32) * N ? int64_t(32) * N : -d64;
				shr_frame<N>(Sm, int(dd >> 5), int(dd & 31), lost);
				stS = lost != 0u;
				stB = 0u;
				sb = c.sk;
				ssm = sx;
				eb = c.exp;
			}
			const bool subtract = ((sb ^ ssm) & SIGN_BIT) != 0u;
			uint32_t R[N];
			uint32_t st;
			int32_t er = eb;
			if (!subtract)
			{
				st = stB | stS;
				uint32_t carry;
				{
					QB_CY_DECL
					QB_ADD_FIRST(R[0], B[0], Sm[0]);
#pragma unroll
					for (int k = 1; k < N; ++k) QB_ADD_NEXT(R[k], B[k], Sm[k]);
					QB_CARRY_OUT(carry);
				}
				if (carry)
				{
					st |= (R[0] & 1u);
#pragma unroll
					for (int k = 0; k < N - 1; ++k) R[k] = (R[k] >> 1) | (R[k + 1] << 31);
					R[N - 1] = (R[N - 1] >> 1) | 0x80000000u;
					er += 1;
				}
			}
			else
			{
				// true value = (B - Sm) + (epsB - epsS), 0 <= eps < 1 unit of the frame's last place.
				// Remainder only on the subtrahend: (B - Sm - 1) + (1 - epsS), done as a borrow-in.
				st = stB | stS;
				const uint32_t bin = (stS && !stB) ? 1u : 0u;
				{
					QB_CY_DECL
					QB_SET_BORROW(bin);
#pragma unroll
					for (int k = 0; k < N; ++k) QB_SUB_NEXT(R[k], B[k], Sm[k]);
				}
				if (stS && stB)
				{
					// net remainder in (-1, 1): decidable when the lowest limb is neither 0 nor all ones and stays a
					// pure sticky limb after normalisation (top limb nonzero: at most 31 leading zero bits)
					if (R[0] == 0u || R[0] == 0xFFFFFFFFu || R[N - 1] == 0u) return false;
					R[0] -= 1u;  // either choice leaves R[0] nonzero: same rounding
				}
				if ((R[N - 1] | R[N - 2]) == 0u)
				{
					uint32_t nz = 0u;
#pragma unroll
					for (int k = 0; k < N; ++k) nz |= R[k];
					if (nz == 0u && st == 0u)
					{
						set_zero(r, 0u);  // exact cancellation: +0 under RNDN
						return true;
					}
					return false;  // deep cancellation: generic path
				}
			}
			return round_frame<L>(r, sb, er, R, st, prec);
		}

		// normalised product frame of two regular numbers: X[0 .. L+3) with the top bit of X[L+2] set, exponent e, st = OR of
		// everything below X[1] (X[0] included); a's limbs are read row by row from a_limbs (stride AS), b's limbs are
		// pulled into registers
		template <int L, int AS>
		QB_HD QB_INLINE void product_frame(uint32_t* X, uint32_t& st, int32_t& e, const uint32_t* a_limbs, int32_t a_exp,
		                                   const mpf<L>& b)
		{
			uint32_t rb[L];
#pragma unroll
			for (int j = 0; j < L; ++j) rb[j] = b.m[j];
			e = a_exp + b.exp;
			if (!mul_top_short<L, AS>(X, st, a_limbs, rb)) mul_top<L, AS>(X, st, a_limbs, rb);
			if ((X[L + 2] & 0x80000000u) == 0u)
			{
#pragma unroll
				for (int k = L + 2; k >= 1; --k) X[k] = (X[k] << 1) | (X[k - 1] >> 31);
				X[0] <<= 1;  // its lowest bit is unknown: the limb is folded into the sticky below
				e -= 1;
			}
			st |= X[0];
		}

		// r = RN(a * b) for REGULAR a given as (sign/kind word, exponent, limbs at stride AS) and regular b
		template <int L, int AS>
		QB_HD QB_INLINE void mul_strided(mpf<L>& r, uint32_t a_sk, int32_t a_exp, const uint32_t* a_limbs, const mpf<L>& b, int prec)
		{
			uint32_t X[L + 3], st;
			int32_t e;
			product_frame<L, AS>(X, st, e, a_limbs, a_exp, b);
			round_frame<L>(r, (a_sk ^ b.sk) & SIGN_BIT, e, X + 1, st, prec);  // top bit set: always succeeds
		}

		// r = RN(a * b)
		template <int L>
		QB_HD QB_INLINE void mul(mpf<L>& r, const mpf<L>& a, const mpf<L>& b, int prec)
		{
			if ((a.sk & 3u) != K_REG || (b.sk & 3u) != K_REG)
			{
				cold_mul(r, a, b, prec);
				return;
			}
			mul_strided<L, 1>(r, a.sk, a.exp, a.m, b, prec);
		}

		// r = RN(a * b + c) for REGULAR a (strided limbs, see mul_strided), regular b and c.  Returns false when the frame
		// cannot decide the rounding (addround); the caller then owns the generic fall-back.
		template <int L, int AS>
		QB_HD QB_INLINE bool fma_strided(mpf<L>& r, uint32_t a_sk, int32_t a_exp, const uint32_t* a_limbs, const mpf<L>& b,
		                                 const mpf<L>& c, int prec)
		{
			uint32_t X[L + 3], st;
			int32_t e;
			product_frame<L, AS>(X, st, e, a_limbs, a_exp, b);
			return addround<L>(r, (a_sk ^ b.sk) & SIGN_BIT, e, X + 1, st, c, prec);
		}

		// r = RN(a * b + c)
		template <int L>
		QB_HD QB_INLINE void fma(mpf<L>& r, const mpf<L>& a, const mpf<L>& b, const mpf<L>& c, int prec)
		{
			if ((a.sk & 3u) != K_REG || (b.sk & 3u) != K_REG || (c.sk & 3u) != K_REG)
			{
				cold_fma(r, a, b, c, prec);
				return;
			}
			if (!fma_strided<L, 1>(r, a.sk, a.exp, a.m, b, c, prec)) cold_fma(r, a, b, c, prec);
		}

		// r = RN(a + b) (negate_b: a - b)
		template <int L>
		QB_HD QB_INLINE void add(mpf<L>& r, const mpf<L>& a, const mpf<L>& b, int prec, bool negate_b = false)
		{
			if ((a.sk & 3u) != K_REG || (b.sk & 3u) != K_REG)
			{
				cold_add(r, a, b, prec, negate_b);
				return;
			}
			uint32_t X[L + 2];
			X[0] = X[1] = 0u;
#pragma unroll
			for (int j = 0; j < L; ++j) X[2 + j] = a.m[j];
			mpf<L> bb;
			copy(bb, b);
			if (negate_b) bb.sk ^= SIGN_BIT;
			if (!addround<L>(r, a.sk, a.exp, X, 0u, bb, prec)) cold_add(r, a, b, prec, negate_b);
		}

		// r = RN(a * v), 0 < |v| < 2^32   (mpfr_mul_ui / mpfr_mul_si with a small multiplier)
		template <int L>
		QB_HD QB_INLINE void mul_small(mpf<L>& r, const mpf<L>& a, int64_t v, int prec)
		{
			const uint64_t u = v < 0 ? uint64_t(0) - uint64_t(v) : uint64_t(v);
			if ((a.sk & 3u) != K_REG || u == 0 || (u >> 32) != 0)
			{
				cold_mul_si(r, a, v, prec);
				return;
			}
			uint32_t X[L + 2];
			uint64_t cy = 0;
			X[0] = 0u;
#pragma unroll
			for (int j = 0; j < L; ++j)
			{
				cy += uint64_t(a.m[j]) * uint32_t(u);
				X[1 + j] = uint32_t(cy);
				cy >>= 32;
			}
			X[L + 1] = uint32_t(cy);
			const uint32_t sgn = (a.sk ^ (v < 0 ? SIGN_BIT : 0u)) & SIGN_BIT;
			if (!round_frame<L>(r, sgn, a.exp + 32, X, 0u, prec)) cold_mul_si(r, a, v, prec);
		}

		// (hi:lo) << t, upper word, 0 <= t <= 32
		QB_HD QB_INLINE uint32_t shl_pair32(uint32_t hi, uint32_t lo, int t)
		{
#if defined(__CUDA_ARCH__)
			return __funnelshift_lc(lo, hi, t);
#else
			return uint32_t((((uint64_t(hi) << 32) | lo) << t) >> 32);
#endif
		}
		// s <- RN(RN(s * n) + c): one step of the reference's Horner evaluation with an integer argument (Polynomial::eval,
		// include/qboot/algebra/polynomial.hpp:66-83: mul_ui, then add, each rounded) as ONE pass: the product is rounded in
		// place as an integer (its dropped bits all sit in the two lowest limbs), shifted once into the normalised frame of
		// the addition and handed to addround -- no intermediate record, no second normalisation.  s, c regular, 0 < n < 2^32.
		// Returns false (s untouched) when addround declines; the caller then takes mul_small + add.
		template <int L>
		QB_HD QB_INLINE bool horner_step(mpf<L>& s, uint32_t n, const mpf<L>& c, int prec)
		{
			constexpr int N = L + 2;
			// T[0 .. L] = s.m * n
			uint32_t T[L + 1];
			{
				uint64_t cy = 0;
#pragma unroll
				for (int j = 0; j < L; ++j)
				{
					cy += uint64_t(s.m[j]) * n;
					T[j] = uint32_t(cy);
					cy >>= 32;
				}
				T[L] = uint32_t(cy);
			}
			// the product has 32 L + nb bits; keep prec of them: round away the k = nb + (32 L - prec) lowest (k <= 63)
			int nb = T[L] ? 32 - clz32(T[L]) : 0;
			const int k = nb + (32 * L - prec);
			if (k > 0)
			{
				uint64_t low = (uint64_t(T[1]) << 32) | T[0];
				const uint64_t mask = (uint64_t(1) << k) - 1u, half = uint64_t(1) << (k - 1);
				const uint64_t rem = low & mask;
				const bool up = rem > half || (rem == half && ((low >> k) & 1u));  // k <= 63: bit k is inside `low`
				low &= ~mask;
				if (up)
				{
					const uint64_t inc = uint64_t(1) << k, sum = low + inc;
					if (sum < inc)
					{
						// the carry leaves the two lowest limbs (probability ~2^-(64 - k))
						bool cyb = true;
#pragma unroll
						for (int j = 2; j <= L; ++j)
						{
							T[j] += cyb ? 1u : 0u;
							cyb = cyb && T[j] == 0u;
						}
						nb = T[L] ? 32 - clz32(T[L]) : 0;
					}
					low = sum;
				}
				T[0] = uint32_t(low);
				T[1] = uint32_t(low >> 32);
			}
			// normalised frame of N = L + 2 limbs: F = T << (64 - nb)
			uint32_t F[N];
			const int t = 32 - nb;  // 0 .. 32
#pragma unroll
			for (int j = 0; j < N; ++j)
			{
				const uint32_t hi = (j >= 1) ? T[j - 1] : 0u, lo = (j >= 2) ? T[j - 2] : 0u;
				F[j] = shl_pair32(hi, lo, t);
			}
			mpf<L> r;
			if (!addround<L>(r, s.sk, s.exp + nb, F, 0u, c, prec)) return false;
			copy(s, r);
			return true;
		}

		// r = RN(a + v) for a small integer v
		template <int L>
		QB_HD QB_INLINE void add_small(mpf<L>& r, const mpf<L>& a, int64_t v, int prec)
		{
			const uint64_t u = v < 0 ? uint64_t(0) - uint64_t(v) : uint64_t(v);
			if ((a.sk & 3u) != K_REG || u == 0 || (u >> 32) != 0)
			{
				cold_add_si(r, a, v, prec);
				return;
			}
			// v as a normalised record (exact: prec >= 64)
			mpf<L> t;
			const int lzv = clz32(uint32_t(u));
#pragma unroll
			for (int j = 0; j < L - 1; ++j) t.m[j] = 0u;
			t.m[L - 1] = uint32_t(u) << lzv;
			t.exp = 32 - lzv;
			t.sk = K_REG | (v < 0 ? SIGN_BIT : 0u);
			uint32_t X[L + 2];
			X[0] = X[1] = 0u;
#pragma unroll
			for (int j = 0; j < L; ++j) X[2 + j] = a.m[j];
			if (!addround<L>(r, a.sk, a.exp, X, 0u, t, prec)) cold_add_si(r, a, v, prec);
		}

		// ---- division ----------------------------------------------------------------------------------------------------
		// reciprocal for 3-by-2 limb division (Moller & Granlund, "Improved division by invariant integers", alg. 6);
		// (d1, d0) normalised (top bit of d1 set)
		QB_HD QB_INLINE uint32_t reciprocal_3by2(uint32_t d1, uint32_t d0)
		{
			uint32_t v = uint32_t(0xFFFFFFFFFFFFFFFFull / d1 - 0x100000000ull);
			uint32_t p = d1 * v + d0;
			if (p < d0)
			{
				--v;
				if (p >= d1)
				{
					--v;
					p -= d1;
				}
				p -= d1;
			}
			uint64_t t = uint64_t(v) * d0;
			uint32_t t1 = uint32_t(t >> 32), t0 = uint32_t(t);
			p += t1;
			if (p < t1)
			{
				--v;
				if (p > d1 || (p == d1 && t0 >= d0)) --v;
			}
			return v;
		}
		// quotient digit of (u2, u1, u0) / (d1, d0) with (u2, u1) < (d1, d0)   (ibid., alg. 5)
		QB_HD QB_INLINE uint32_t div_3by2(uint32_t u2, uint32_t u1, uint32_t u0, uint32_t d1, uint32_t d0, uint32_t v)
		{
			uint64_t q = uint64_t(v) * u2 + ((uint64_t(u2) << 32) | u1);
			uint32_t q1 = uint32_t(q >> 32), q0 = uint32_t(q);
			uint32_t r1 = u1 - q1 * d1;
			uint64_t rr = ((uint64_t(r1) << 32) | u0) - uint64_t(d0) * q1 - ((uint64_t(d1) << 32) | d0);
			q1 += 1u;
			if (uint32_t(rr >> 32) >= q0)
			{
				q1 -= 1u;
				rr += (uint64_t(d1) << 32) | d0;
			}
			if (rr >= ((uint64_t(d1) << 32) | d0)) q1 += 1u;
			return q1;
		}

		// r = RN(a / b): schoolbook long division with the remainder window, the divisor and the quotient in registers
		// (static indices only): per quotient limb one 3-by-2 estimate, L carry-free IMAD.WIDE, two subtract-with-borrow
		// chains and a rare add-back.
		template <int L>
		QB_HD QB_INLINE void div(mpf<L>& r, const mpf<L>& a, const mpf<L>& b, int prec)
		{
			if ((a.sk & 3u) != K_REG || (b.sk & 3u) != K_REG)
			{
				cold_div(r, a, b, prec);
				return;
			}
			constexpr int LP = (L + 1) & ~1;
			uint32_t D[LP], R[L + 1], Q[L + 1];
#pragma unroll
			for (int i = 0; i < LP; ++i) D[i] = i < L ? b.m[i] : 0u;
			// numerator below the divisor: a.m, or a.m / 2 when a.m >= b.m
			bool ge = true, decided = false;
#pragma unroll
			for (int i = L - 1; i >= 0; --i)
				if (!decided && a.m[i] != D[i])
				{
					ge = a.m[i] > D[i];
					decided = true;
				}
			int32_t e = a.exp - b.exp;
			uint32_t nextlimb = 0u;
			if (ge)
			{
#pragma unroll
				for (int i = 0; i < L; ++i) R[i] = (a.m[i] >> 1) | ((i + 1 < L ? a.m[i + 1] : 0u) << 31);
				nextlimb = a.m[0] << 31;
				e += 1;
			}
			else
			{
#pragma unroll
				for (int i = 0; i < L; ++i) R[i] = a.m[i];
			}
			R[L] = 0u;
#pragma unroll
			for (int i = 0; i <= L; ++i) Q[i] = 0u;
			const uint32_t d1 = D[L - 1], d0 = D[L - 2];
			const uint32_t v = reciprocal_3by2(d1, d0);
#pragma unroll 1
			for (int j = 0; j <= L; ++j)
			{
				// bring in the next numerator limb
#pragma unroll
				for (int i = L; i >= 1; --i) R[i] = R[i - 1];
				R[0] = nextlimb;
				nextlimb = 0u;
				uint32_t qhat;
				if (R[L] > d1 || (R[L] == d1 && R[L - 1] >= d0))
					qhat = 0xFFFFFFFFu;
				else
					qhat = div_3by2(R[L], R[L - 1], R[L - 2], d1, d0, v);
				// R -= qhat * D: even and odd limb products form carry-free 64-bit pairs
				uint32_t TE[LP + 1], TO[LP + 1];
#pragma unroll
				for (int k = 0; k < LP; k += 2)
				{
					uint64_t te = uint64_t(qhat) * D[k];
					uint64_t to = uint64_t(qhat) * D[k + 1];
					TE[k] = uint32_t(te);
					TE[k + 1] = uint32_t(te >> 32);
					TO[k] = uint32_t(to);       // position k + 1
					TO[k + 1] = uint32_t(to >> 32);
				}
				uint32_t bo1, bo2;
				{
					QB_CY_DECL
					QB_SUB_FIRST(R[0], R[0], TE[0]);
#pragma unroll
					for (int k = 1; k <= L; ++k) QB_SUB_NEXT(R[k], R[k], (k < LP ? TE[k] : 0u));
					QB_BORROW_OUT(bo1);
				}
				{
					QB_CY_DECL
					QB_SUB_FIRST(R[1], R[1], TO[0]);
#pragma unroll
					for (int k = 2; k <= L; ++k) QB_SUB_NEXT(R[k], R[k], (k - 1 < LP ? TO[k - 1] : 0u));
					QB_BORROW_OUT(bo2);
				}
				// limbs of qhat * D above position L are zero except when LP > L (padding limb of D is zero): nothing lost
				if (bo1 | bo2)
				{
					// estimate one too large: add the divisor back
					qhat -= 1u;
					QB_CY_DECL
					QB_ADD_FIRST(R[0], R[0], D[0]);
#pragma unroll
					for (int k = 1; k <= L; ++k) QB_ADD_NEXT(R[k], R[k], (k < L ? D[k] : 0u));
				}
#pragma unroll
				for (int i = L; i >= 1; --i) Q[i] = Q[i - 1];
				Q[0] = qhat;
			}
			uint32_t st = 0u;
#pragma unroll
			for (int i = 0; i <= L; ++i) st |= R[i];
			uint32_t X[L + 2];
			X[0] = 0u;
#pragma unroll
			for (int i = 0; i <= L; ++i) X[1 + i] = Q[i];
			if (!round_frame<L>(r, (a.sk ^ b.sk) & SIGN_BIT, e, X, st, prec)) cold_div(r, a, b, prec);
		}
	}  // namespace fast
}  // namespace qb
