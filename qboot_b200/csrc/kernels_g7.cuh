// qboot_b200 -- k_series_coop: the series recurrence of a SMALL batch (one PMP: 1-10 k series jobs), eight warps per 32 jobs.
//
//     b[n] = -(sum_{i=1}^{min(n,K-1)} p_i(n) b[n-i]) / p_0(n)                          (src/hor_recursion.cpp:31-37)
//
// k_series (kernels_g1.cuh) gives every job one thread: 400 dependent steps of ~48 k instructions each, which a lone warp
// per scheduler walks in 47 ms however few jobs there are.  A step has far more parallelism than that: the polynomials do
// not depend on b at all, the products p_i(n) b[n-i] with i >= 2 do not depend on b[n-1], and only the rounded additions
// and the division form the chain.  Here a CTA of eight warps owns 32 jobs (lane = job) and every warp has ONE role, so a
// warp never diverges:
//     warp 0       the chain: s = RN(X_1), s = RN(s + X_i) for i = 2.., b[n] = RN(-s / p_0(n))
//     warp 1       X_1 = p_1(n) b[n-1] as soon as b[n-1] exists; p_1(n+1) and p_0(n+1) while warp 0 works
//     warps 2..7   p_w(n+1) and X_w = p_w(n+1) b[n+1-w], one step ahead of the chain
// X_i is the EXACT product as a normalised frame (fast::product_frame), handed over in shared memory; the additions use
// the same fast::round_frame / fast::addround as k_series and the same generic fall-backs, in the same order, so b is
// the same bit for bit.  Two CTA barriers per step.  Critical path per step: product + 6 additions + division
// (~12 k instructions) instead of ~48 k.
#pragma once

#include "kernels_g1.cuh"

namespace qb
{
	constexpr int QB_COOP_JOBS = 32;     // jobs per CTA (one per lane)
	constexpr int QB_COOP_WARPS = 8;     // roles
	constexpr int QB_COOP_T = QB_COOP_JOBS * QB_COOP_WARPS;
#ifndef QB_COOP_CTAS
#define QB_COOP_CTAS 1
#endif
	// shared memory, in words; every array is [word][lane] so that a warp's access is one conflict-free row
	template <int L>
	struct CoopSmem
	{
		static constexpr int FRAME = L + 5;                                   // header, sticky, exponent, L + 2 limbs
		static constexpr int P_SLOTS = 9;                                     // p_0 (two generations), p_1 .. p_7
		static constexpr int off_p = 0;                                       // [P_SLOTS][L][32]
		static constexpr int off_meta = off_p + P_SLOTS * L * 32;             // [P_SLOTS][2][32]: sign/kind word, exponent
		static constexpr int off_x1 = off_meta + P_SLOTS * 2 * 32;            // [FRAME][32]
		static constexpr int off_xn = off_x1 + FRAME * 32;                    // [2][6][FRAME][32]
		static constexpr int words = off_xn + 2 * 6 * FRAME * 32;
		static constexpr size_t bytes = size_t(words) * sizeof(uint32_t);
	};

	// p = p_ip(n), exactly as k_series evaluates it: in the fixed-point frame when that is exact, in floating point otherwise
	template <int L>
	__device__ __forceinline__ void coop_poly(mpf<L>& p, uint32_t n, int ip, int deg, int cj, int n_jobs, const uint32_t* __restrict__ cw,
	                                          const uint32_t* __restrict__ coefx, int prec)
	{
		constexpr int W = L + 2, CS = fx_words<L>() * 32;
		const int st = deg + 1;
		bool fixed = false;
		const uint32_t* q = coefx + coefx_offset<L>(cj, ip * st + deg);
		uint32_t sg = q[0];
		if (sg & 1u)
		{
			const int32_t E = int32_t(q[32]);
			uint32_t V[W], C[W];
#pragma unroll
			for (int j = 0; j < W; ++j) V[j] = q[(2 + j) * 32];
			fixed = true;
			QB_ROLL
			for (int d = deg - 1; d >= 0; --d)
			{
				q -= CS;
				const uint32_t csg = q[0];
#pragma unroll
				for (int j = 0; j < W; ++j) C[j] = q[(2 + j) * 32];
				if (fixed) fixed = fx::horner_step<L>(V, sg, n, C, csg, prec);
			}
			if (fixed) fixed = fx::to_record<L>(p, V, sg, E);
		}
		if (!fixed)
		{
			mpf<L> cd;
			coef_load(p, cw, ip * st + deg, cj, n_jobs);
			QB_ROLL
			for (int d = deg - 1; d >= 0; --d)
			{
				coef_load(cd, cw, ip * st + d, cj, n_jobs);
				if ((p.sk & 3u) != K_REG || (cd.sk & 3u) != K_REG || !fast::horner_step(p, n, cd, prec))
				{
					fast::mul_small(p, p, int64_t(n), prec);
					fast::add(p, p, cd, prec);
				}
			}
		}
	}

	// X = p * bp as a frame in shared memory (header bit 0 = valid, bit 31 = sign); invalid when an operand is not regular
	template <int L>
	__device__ __forceinline__ void coop_product(uint32_t* frame, const uint32_t* pl, uint32_t psk, int32_t pexp, const mpf<L>& bp)
	{
		if ((psk & 3u) != K_REG || (bp.sk & 3u) != K_REG)
		{
			frame[0] = 0u;
			return;
		}
		uint32_t X[L + 3], stt;
		int32_t e;
		fast::product_frame<L, 32>(X, stt, e, pl, pexp, bp);
		frame[0] = 1u | ((psk ^ bp.sk) & SIGN_BIT);
		frame[32] = stt;
		frame[64] = uint32_t(e);
#pragma unroll
		for (int j = 0; j < L + 2; ++j) frame[(3 + j) * 32] = X[1 + j];
	}

	template <int L>
	__global__ void __launch_bounds__(QB_COOP_T, QB_COOP_CTAS)
	    k_series_coop(DevCtx cx, int n_jobs, const mpf<L>* __restrict__ delta, const uint32_t* __restrict__ spin, const mpf<L>* __restrict__ b0,
	                  const mpf<L>* __restrict__ coef, const uint32_t* __restrict__ coefx, mpf<L>* b)
	{
		using S = CoopSmem<L>;
		extern __shared__ uint32_t sh[];
		const int w = int(threadIdx.x >> 5), lane = int(threadIdx.x & 31u);
		const int job = blockIdx.x * QB_COOP_JOBS + lane;
		const int prec = cx.prec;
		const uint32_t n_max = cx.n_max;
		bool active = job < n_jobs;
		const int cj = active ? job : 0;
		mpf<L>* bj = b + size_t(cj) * (n_max + 1);
		if (active && is_zero(delta[job]))
		{
			// unit operator: b = (b0, 0, 0, ...)                                      (src/hor_recursion.cpp:20)
			if (w == 0)
			{
				copy(bj[0], b0[spin[job]]);
				for (uint32_t n = 1; n <= n_max; ++n) set_zero(bj[n]);
			}
			active = false;
		}
		if (active && w == 0) copy(bj[0], b0[spin[job]]);
		int K = 8, deg = 4;
		if (active && spin[job] == 0)
		{
			K = 6;
			deg = 3;
		}
		asm volatile("" : "+r"(K), "+r"(deg));  // one copy of every loop body
		const uint32_t* cw = reinterpret_cast<const uint32_t*>(coef);
		uint32_t* const p_base = sh + S::off_p + lane;
		uint32_t* const meta = sh + S::off_meta + lane;
		uint32_t* const x1 = sh + S::off_x1 + lane;
		uint32_t* const xn = sh + S::off_xn + lane;
		// slot of p_i: i >= 1 -> 1 + i; p_0 of step n -> n & 1
		const auto p_limbs = [&](int slot) { return p_base + slot * L * 32; };
		const auto put_p = [&](int slot, const mpf<L>& p) {
			uint32_t* d = p_limbs(slot);
#pragma unroll
			for (int j = 0; j < L; ++j) d[j * 32] = p.m[j];
			meta[(slot * 2) * 32] = p.sk;
			meta[(slot * 2 + 1) * 32] = uint32_t(p.exp);
		};
		// before the first step: p_1(1) and p_0(1)
		if (w == 1 && active)
		{
			mpf<L> p;
			coop_poly<L>(p, 1u, 1, deg, cj, n_jobs, cw, coefx, prec);
			put_p(2, p);
			coop_poly<L>(p, 1u, 0, deg, cj, n_jobs, cw, coefx, prec);
			put_p(1, p);
		}
		__syncthreads();
		for (uint32_t n = 1; n <= n_max; ++n)
		{
			// ---- first half: X_1(n) needs b[n-1]; warps 2..7 evaluate their polynomial of step n + 1
			if (w == 1)
			{
				if (active)
				{
					mpf<L> bp;
					ldg_rec64(bp, &bj[n - 1u]);
					coop_product<L>(x1, p_limbs(2), meta[4 * 32], int32_t(meta[5 * 32]), bp);
				}
			}
			else if (w >= 2)
			{
				const uint32_t m = n + 1u;
				if (active && w < K && uint32_t(w) <= m && m <= n_max)
				{
					mpf<L> p;
					coop_poly<L>(p, m, w, deg, cj, n_jobs, cw, coefx, prec);
					put_p(1 + w, p);
				}
			}
			__syncthreads();
			// ---- second half: the chain; meanwhile the products of step n + 1 and warp 1's polynomials
			if (w == 0)
			{
				if (active)
				{
					mpf<L> sum;
					set_zero(sum, 0u);
					QB_ROLL
					for (int i = 1; i <= 7; ++i)
					{
						if (i >= K || uint32_t(i) > n) continue;
						const uint32_t* f = i == 1 ? x1 : xn + ((int(n & 1u) * 6 + (i - 2)) * S::FRAME) * 32;
						const uint32_t hdr = f[0];
						bool done = false;
						mpf<L> r;
						if ((hdr & 1u) && (i == 1 || (sum.sk & 3u) == K_REG))
						{
							uint32_t X[L + 2];
#pragma unroll
							for (int j = 0; j < L + 2; ++j) X[j] = f[(3 + j) * 32];
							const uint32_t stt = f[32], sgn = hdr & SIGN_BIT;
							const int32_t e = int32_t(f[64]);
							if (i == 1)
								done = fast::round_frame<L>(r, sgn, e, X, stt, prec);  // fma(+0, p_1 b[n-1]) rounds like the product
							else
								done = fast::addround<L>(r, sgn, e, X, stt, sum, prec);
						}
						if (!done)
						{
							// special values, deep cancellation: the generic routines on the operands themselves
							mpf<L> pp, bb;
							coop_poly<L>(pp, n, i, deg, cj, n_jobs, cw, coefx, prec);
							ldg_rec64(bb, &bj[n - uint32_t(i)]);
							if (i == 1)
								fast::cold_mul(r, pp, bb, prec);
							else
								fast::cold_fma(r, pp, bb, sum, prec);
						}
						copy(sum, r);
					}
					// b[n] = -sum / p_0(n)
					mpf<L> p0, bn;
					const int s0 = int(n & 1u);
					series_load<L, 32>(p0, p_limbs(s0), meta[(s0 * 2) * 32], int32_t(meta[(s0 * 2 + 1) * 32]));
					neg(sum);
					fast::div(bn, sum, p0, prec);
					copy(bj[n], bn);
				}
			}
			else if (w == 1)
			{
				const uint32_t m = n + 1u;
				if (active && m <= n_max)
				{
					mpf<L> p;
					coop_poly<L>(p, m, 1, deg, cj, n_jobs, cw, coefx, prec);
					put_p(2, p);
					coop_poly<L>(p, m, 0, deg, cj, n_jobs, cw, coefx, prec);
					put_p(int(m & 1u), p);
				}
			}
			else
			{
				const uint32_t m = n + 1u;
				if (active && w < K && uint32_t(w) <= m && m <= n_max)
				{
					mpf<L> bp;
					ldg_rec64(bp, &bj[m - uint32_t(w)]);
					const int slot = 1 + w;
					coop_product<L>(xn + ((int(m & 1u) * 6 + (w - 2)) * S::FRAME) * 32, p_limbs(slot), meta[(slot * 2) * 32],
					                int32_t(meta[(slot * 2 + 1) * 32]), bp);
				}
			}
			__syncthreads();
		}
	}
}  // namespace qb
