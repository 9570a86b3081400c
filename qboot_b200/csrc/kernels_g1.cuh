// qboot_b200 -- CUDA kernels of the block-table pipeline (sm_100a), templated on the limb count L; split by kernel
// group so that each translation unit (kernel_inst.cu -DQB_KGROUP=g) only parses what it instantiates:
//   kernels_g0.cuh  k_rec_coeffs  one thread per (series job, recurrence coefficient)              -> coef[job][40]
//                   k_polyvals    one thread per (series job, n): the recurrence polynomials at n  -> pv[job][n][8]
//                   k_ops         diagnostic: one primitive elementwise
//   kernels_g1.cuh  k_series      one thread per series job, n_max sequential steps                -> b[job][0..n_max]
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

namespace qb
{
	// b[n] = -(sum_{i=1}^{min(n,K-1)} p_i(n) b[n-i]) / p_0(n), one thread per series job (src/hor_recursion.cpp:31-37).
	// The running sum stays in registers; p_i(n) (from k_polyvals) and the previous terms are read where they lie in HBM.
	template <int L>
	__global__ void k_series(DevCtx cx, int n_jobs, const mpf<L>* __restrict__ delta, const uint32_t* __restrict__ spin,
	                         const mpf<L>* __restrict__ b0, const mpf<L>* __restrict__ pv, mpf<L>* __restrict__ b)
	{
		int job = blockIdx.x * blockDim.x + threadIdx.x;
		if (job >= n_jobs) return;
		mpf<L>* bj = b + size_t(job) * (cx.n_max + 1);
		copy(bj[0], b0[job]);
		if (is_zero(delta[job]))
		{
			// unit operator: b = (b0, 0, 0, ...)                                  (src/hor_recursion.cpp:20)
			for (uint32_t n = 1; n <= cx.n_max; ++n) set_zero(bj[n]);
			return;
		}
		const int K = spin[job] == 0 ? 6 : 8;
		const int prec = cx.prec;
		const mpf<L>* pj = pv + size_t(job) * cx.n_max * 8;
		for (uint32_t n = 1; n <= cx.n_max; ++n)
		{
			const mpf<L>* p = pj + size_t(n - 1) * 8;
#ifdef __CUDA_ARCH__
			// the next step's polynomial values (8 contiguous records per lane) are pulled into L1 while this step computes
			if (n < cx.n_max)
			{
				const char* q = reinterpret_cast<const char*>(p + 8);
#pragma unroll
				for (int o = 0; o < int(8 * sizeof(mpf<L>)); o += 128) asm volatile("prefetch.global.L1 [%0];" ::"l"(q + o));
			}
#endif
			mpf<L> sum, bn;
			fast::mul(sum, p[1], bj[n - 1], prec);  // fma(0, p_1 b[n-1]) rounds like the product
			for (int i = 2; i < K; ++i)
				if (uint32_t(i) <= n) fast::fma(sum, p[i], bj[n - uint32_t(i)], sum, prec);
			neg(sum);
			fast::div(bn, sum, p[0], prec);
			copy(bj[n], bn);
		}
	}

}  // namespace qb
