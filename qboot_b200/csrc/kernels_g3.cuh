// qboot_b200 -- CUDA kernels of the block-table pipeline (sm_100a), templated on the limb count L; split by kernel
// group so that each translation unit (kernel_inst.cu -DQB_KGROUP=g) only parses what it instantiates:
//   kernels_g0.cuh  k_rec_coeffs  one thread per (series job, recurrence coefficient)              -> coef[job][40]
//                   k_polyvals    one thread per (series job, n): the recurrence polynomials at n  -> pv[job][n][8]
//                   k_ops         diagnostic: one primitive elementwise
//   kernels_g1.cuh  k_series      one thread per series job, n_max sequential steps                -> b[job][0..n_max]
//   kernels_g2.cuh  k_scale       one thread per (table, n): the lambda+1 scaled coefficient arrays -> a[table][k][n]
//                   k_horner      one thread per (table, k): Horner sum, times rho^q / k!           -> fr[table][k]
//   kernels_g3.cuh  k_finish      one thread per table: rho->z matrix and the transverse recursion  -> g[table][dim_M]
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
	template <int L>
	__global__ void k_finish(DevCtx cx, int n_tables, const mpf<L>* __restrict__ delta, const uint32_t* __restrict__ spin,
	                         const mpf<L>* __restrict__ S, const mpf<L>* __restrict__ P, const mpf<L>* __restrict__ fr,
	                         mpf<L>* __restrict__ g)
	{
		int t = blockIdx.x * blockDim.x + threadIdx.x;
		if (t >= n_tables) return;
		const int lambda = int(cx.lambda), nk = lambda + 1, dimM = cf_dim(lambda);
		mpf<L>* f = g + size_t(t) * dimM;
		const mpf<L>* M = reinterpret_cast<const mpf<L>*>(cx.mat);
		mpf<L> z;
		for (int i = 0; i <= lambda; ++i)
		{
			rho_to_z_entry(z, fr + size_t(t) * nk, 1, M, lambda, i, cx.prec);
			copy(f[i], z);
		}
		expand_off_diag(f, 1, lambda, delta[t], spin[t], S[t], P[t], cx.eps, cx.prec);
	}

}  // namespace qb
