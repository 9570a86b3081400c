// qboot_b200 -- CUDA kernels of the block-table pipeline (sm_100a), templated on the limb count L; split by kernel
// group so that each translation unit (kernel_inst.cu -DQB_KGROUP=g) only parses what it instantiates:
//   kernels_g0.cuh  k_rec_coeffs  one thread per (series job, recurrence coefficient)              -> coef, word-planar across jobs
//                   k_ops         diagnostic: one primitive elementwise
//   kernels_g1.cuh  k_coef_fx     one thread per (polynomial, job): fixed-point frames of the recurrence polynomials -> coefx
//                   k_series      one thread per series job over a front group: recurrence polynomials + the chain   -> b[job][0..n_max]
//   kernels_g7.cuh  k_series_coop the same recurrence for small batches: eight warps (roles) per 32 jobs                   -> b
//   kernels_g2.cuh  k_scale       one thread per (table, n): the lambda+1 scaled coefficient arrays -> a[table][k][n]
//                   k_horner      one thread per (table, k): Horner sum, times rho^q / k!           -> fr[table][k]
//   kernels_g3.cuh  k_rho_to_z, k_offdiag_val, k_offdiag_close: rho->z matrix and the transverse recursion, level by level -> g[table][dim_M]
//   kernels_g4.cuh  k_pow         4^Delta and rho^(Delta-k);  k_vtod, k_fblock: v^d tables and F/H blocks
// Numbers are records of L + 2 words (mpf<L>), arrays of records in HBM.  Arithmetic: mpf.cuh / fastops.cuh / hor.cuh /
// block.cuh, all force-inlined (see mpf.cuh for why).
#pragma once

#include <cuda_runtime.h>

#define QB_TABLE_QUAL static __device__
#include "block.cuh"
#include "launch.cuh"
#include "polyfx.cuh"

namespace qb
{
	template <int L>
	__device__ __forceinline__ void ldg_rec64(mpf<L>& d, const mpf<L>* src)
	{
		if constexpr (L % 2 == 0)
		{
			const uint2* sp = reinterpret_cast<const uint2*>(src);
			uint2 v = sp[0];
			d.sk = v.x;
			d.exp = int32_t(v.y);
#pragma unroll
			for (int j = 0; j < L / 2; ++j)
			{
				v = sp[1 + j];
				d.m[2 * j] = v.x;
				d.m[2 * j + 1] = v.y;
			}
		}
		else
			copy(d, *src);
	}

	// Fixed-point frames of the recurrence polynomials (polyfx.cuh): one thread per (polynomial, job), job fastest.  Chooses
	// the frame exponent from the coefficients' exponents and n_max, converts every coefficient and records whether all of
	// them are exact in the frame; k_series evaluates such polynomials on integers and the others in floating point.
	template <int L>
	__global__ void k_coef_fx(DevCtx cx, int n_jobs, const uint32_t* __restrict__ spin, const mpf<L>* __restrict__ coef,
	                          uint32_t* __restrict__ coefx)
	{
		constexpr int W = L + 2, FW = fx_words<L>();
		const int gid = blockIdx.x * blockDim.x + threadIdx.x;
		const int i = gid / n_jobs, job = gid - i * n_jobs;
		if (i >= 8) return;
		const int K = spin[job] == 0 ? 6 : 8, deg = spin[job] == 0 ? 3 : 4, st = deg + 1;
		if (i >= K) return;
		const uint32_t* cw = reinterpret_cast<const uint32_t*>(coef);
		int32_t e[5];
		bool have[5];
		for (int d = 0; d <= deg; ++d)
		{
			const uint32_t* p = cw + size_t(i * st + d) * W * size_t(n_jobs) + job;
			have[d] = (p[0] & 3u) == K_REG;
			e[d] = int32_t(p[size_t(n_jobs)]);
		}
		const int nbits = 32 - clz32(cx.n_max);
		const int32_t E = fx::frame_exponent(e, have, deg, nbits);
		bool ok = true;
		for (int d = 0; d <= deg; ++d)
		{
			mpf<L> c;
			coef_load(c, cw, i * st + d, job, n_jobs);
			uint32_t V[W], sg;
			ok = fx::from_record<L>(V, sg, c, E) && ok;
			uint32_t* q = coefx + coefx_offset<L>(job, i * st + d);
			q[0] = sg;
			q[32] = uint32_t(E);
			for (int j = 0; j < W; ++j) q[(2 + j) * 32] = V[j];
		}
		// the verdict goes into the leading coefficient's header
		coefx[coefx_offset<L>(job, i * st + deg)] |= ok ? 1u : 0u;
	}

	// The series recurrence b[n] = -(sum_{i=1}^{min(n,K-1)} p_i(n) b[n-i]) / p_0(n)       (src/hor_recursion.cpp:31-37),
	// one thread per series job, launched over ALL jobs of a batch at once: the recurrence is a chain of n_max dependent
	// steps, so the only parallelism is across jobs, and the limb-product pipe needs >= 4 resident warps per scheduler to
	// stay busy (a chunk-sized launch left most schedulers with no warp at all: 8 % of the limb-product peak).
	//   * the recurrence polynomials p_i(n) are evaluated in the thread, right before they are used, by Horner with the
	//     integer n exactly as the reference's Polynomial::eval does (include/qboot/algebra/polynomial.hpp:66-83: mul_ui,
	//     add, each rounded): they never travel through HBM (a separate k_polyvals kernel used to write 435 KB per table);
	//   * p_i(n) -- the row operand of the product -- and the running sum live in SHARED memory, limb-major and
	//     conflict-free (limb j of thread x at sh[j * 128 + x]), like the chain operands of k_scale / k_horner: the rolled
	//     product loop indexes its row operand dynamically, and thread-local records of 75 k resident threads overflow L1;
	//   * one copy of the product code serves the first term (a plain rounded product) and the fused multiply-adds.
// One CTA of 384 threads per SM (<= 170 registers): measured best of 128 x 3, 192 x 2, 256 x 1, 256 x 2, 320 x 1, 384 x 1,
	// 448 x 1, 512 x 1 -- the CTA is the lockstep group, so one large CTA shares instruction fetches best, and 128-register
	// variants spill 3.4 KB per thread (profiles/r2_summary.md)
#ifndef QB_SERIES_T
#define QB_SERIES_T 384
#endif
	constexpr int QB_SERIES_THREADS = QB_SERIES_T;
	// small batches (a single PMP: 1-6 k jobs) run CTAs of 128 threads instead: one warp per scheduler on as many SMs as the
	// batch reaches -- a step of a lone warp takes a third of a step among twelve
	constexpr int QB_SERIES_THREADS_SMALL = 128;
	template <int L, int T>
	constexpr size_t series_smem_bytes()
	{
		return size_t(2) * L * T * sizeof(uint32_t);
	}
	template <int L, int T>
	__device__ __forceinline__ void series_store(uint32_t* my, uint32_t& sk, int32_t& exp, const mpf<L>& v)
	{
		sk = v.sk;
		exp = v.exp;
#pragma unroll
		for (int j = 0; j < L; ++j) my[j * T] = v.m[j];
	}
	template <int L, int T>
	__device__ __forceinline__ void series_load(mpf<L>& v, const uint32_t* my, uint32_t sk, int32_t exp)
	{
		v.sk = sk;
		v.exp = exp;
#pragma unroll
		for (int j = 0; j < L; ++j) v.m[j] = my[j * T];
	}
#ifndef QB_SERIES_CTAS
#define QB_SERIES_CTAS 1
#endif
	template <int L, int T>
	__global__ void __launch_bounds__(T, QB_SERIES_CTAS)
	    k_series(DevCtx cx, int n_jobs, const mpf<L>* __restrict__ delta, const uint32_t* __restrict__ spin, const mpf<L>* __restrict__ b0,
	             const mpf<L>* __restrict__ coef, const uint32_t* __restrict__ coefx, mpf<L>* __restrict__ b)
	{
		extern __shared__ uint32_t sh[];
		uint32_t* myp = sh + threadIdx.x;                           // p_i(n)
		uint32_t* mys = sh + L * T + threadIdx.x;   // running sum
		const int job = blockIdx.x * T + threadIdx.x;
		const int prec = cx.prec;
		const uint32_t n_max = cx.n_max;
		bool active = job < n_jobs;
		mpf<L>* bj = b + size_t(active ? job : 0) * (n_max + 1);
		if (active)
		{
			copy(bj[0], b0[spin[job]]);  // b[0] by spin (a table of at most 401 records)
			if (is_zero(delta[job]))
			{
				// unit operator: b = (b0, 0, 0, ...)                                  (src/hor_recursion.cpp:20)
				for (uint32_t n = 1; n <= n_max; ++n) set_zero(bj[n]);
				active = false;
			}
		}
		int K = 8, deg = 4;
		if (active && spin[job] == 0)
		{
			K = 6;
			deg = 3;
		}
		// keep ONE copy of the loop bodies below (no per-spin clones: the kernel must stay inside the instruction cache)
		asm volatile("" : "+r"(K), "+r"(deg));
		const int st = deg + 1;
		const uint32_t* cw = reinterpret_cast<const uint32_t*>(coef);  // word-planar across the jobs of this launch (launch.cuh)
		const int cj = active ? job : 0;
		uint32_t psk, ssk = K_ZERO;
		int32_t pexp, sexp = 0;
		for (uint32_t n = 1; n <= n_max; ++n)
		{
			// i = 1 .. 7: the terms of the sum (those with i < K, i <= n); i = 8: the divisor p_0(n).
			// The warps of the CTA walk the body together (a barrier per term): one pass over the ~5 k instructions of a term
			// costs an instruction fetch per SM instead of one per warp -- unsynchronised, this kernel spent 9 of 10 issue
			// slots waiting for instructions.
			QB_ROLL
			for (int i = 1; i <= 8; ++i)
			{
				__syncthreads();
				const bool divisor = i == 8;
				if (!active || (!divisor && (i >= K || uint32_t(i) > n))) continue;
				const int ip = divisor ? 0 : i;
				mpf<L> p;
				// the polynomial in its fixed-point frame (polyfx.cuh) when its coefficients are exact there ...
				bool fixed = false;
				{
					constexpr int W = L + 2, FW = fx_words<L>();
					constexpr int CS = FW * 32;  // words from one coefficient to the next
					// The coefficients (5.8 KB per job) stream from DRAM at every step -- the jobs in flight hold 300 MB of
					// them.  Each lane touches one line of the NEXT polynomial's coefficients now, a polynomial ahead of their use
					// (the lanes of a warp read the same lines: block layout, launch.cuh), so that the loads below find them in L2.
					{
						const int ipn = divisor ? 1 : (i + 1 < K ? i + 1 : 0);
						const uint32_t* nb = coefx + coefx_offset<L>(cj & ~31, ipn * st);
						QB_ROLL
						for (int idx = int(threadIdx.x & 31u); idx < st * FW; idx += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(nb + idx * 32));
					}
					const uint32_t* q = coefx + coefx_offset<L>(cj, ip * st + deg);
					uint32_t sg = q[0];
					if (sg & 1u)
					{
						const int32_t E = int32_t(q[32]);
						uint32_t V[W], C[W], Cn[W], csgn;
#pragma unroll
						for (int j = 0; j < W; ++j) V[j] = q[(2 + j) * 32];
						// software pipeline: the coefficient of the next Horner step is in flight while this one is computed
						// (without it the series stage is 3 % slower although the loop is 34 moves shorter)
						q -= CS;
						csgn = q[0];
#pragma unroll
						for (int j = 0; j < W; ++j) Cn[j] = q[(2 + j) * 32];
						fixed = true;
						QB_ROLL
						for (int d = deg - 1; d >= 0; --d)
						{
							const uint32_t csg = csgn;
#pragma unroll
							for (int j = 0; j < W; ++j) C[j] = Cn[j];
							if (d > 0)
							{
								q -= CS;
								csgn = q[0];
#pragma unroll
								for (int j = 0; j < W; ++j) Cn[j] = q[(2 + j) * 32];
							}
							if (fixed) fixed = fx::horner_step<L>(V, sg, n, C, csg, prec);
						}
						if (fixed) fixed = fx::to_record<L>(p, V, sg, E);
					}
				}
				// ... and in floating point otherwise (frame not exact, a deep cancellation, a zero)
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
				// p_i(n) goes to its shared-memory slot at once: only its sign / kind word and exponent stay in registers
				series_store<L, T>(myp, psk, pexp, p);
				if (divisor)
				{
					// b[n] = -sum / p_0(n)
					mpf<L> sum, bn, p0;
					series_load<L, T>(sum, mys, ssk, sexp);
					series_load<L, T>(p0, myp, psk, pexp);
					neg(sum);
					fast::div(bn, sum, p0, prec);
					copy(bj[n], bn);
					continue;
				}
				mpf<L> bp;
				ldg_rec64(bp, &bj[n - uint32_t(i)]);
				bool done = false;
				if ((psk & 3u) == K_REG && (bp.sk & 3u) == K_REG && (i == 1 || (ssk & 3u) == K_REG))
				{
					uint32_t X[L + 3], stt;
					int32_t e;
					fast::product_frame<L, T>(X, stt, e, myp, pexp, bp);
					const uint32_t sgn = (psk ^ bp.sk) & SIGN_BIT;
					mpf<L> r;
					if (i == 1)
						done = fast::round_frame<L>(r, sgn, e, X + 1, stt, prec);  // fma(+0, p_1 b[n-1]) rounds like the product
					else
					{
						mpf<L> sum;
						series_load<L, T>(sum, mys, ssk, sexp);
						done = fast::addround<L>(r, sgn, e, X + 1, stt, sum, prec);
					}
					if (done) series_store<L, T>(mys, ssk, sexp, r);
				}
				if (!done)
				{
					// special values, deep cancellation: the generic routines (same correctly rounded result)
					mpf<L> sum, r, pp, bb;
					series_load<L, T>(pp, myp, psk, pexp);
					ldg_rec64(bb, &bj[n - uint32_t(i)]);
					if (i == 1)
						fast::cold_mul(r, pp, bb, prec);
					else
					{
						series_load<L, T>(sum, mys, ssk, sexp);
						fast::cold_fma(r, pp, bb, sum, prec);
					}
					series_store<L, T>(mys, ssk, sexp, r);
				}
			}
		}
	}

}  // namespace qb
