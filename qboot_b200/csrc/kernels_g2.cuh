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
	// rho-derivative stage, part 1: every scaled coefficient array a_k[n], k = 0..lambda (reference: h *= 4^Delta,
	// src/hor_recursion.cpp:43, then lambda times RealFunctionWithPower::derivate, real_function.hpp:236-240):
	//   a_0[n] = b[n] * 4^Delta,   a_k[n] = a_{k-1}[n] * (q_{k-1} + n),   q_k = q_{k-1} - 1,  q_0 = Delta
	// One thread per (table, n) walks k: no dependency between threads, no barrier.  Layout a[table][k][n] (n fastest) so
	// that part 2 streams each chain contiguously.  The arrays (lambda+1 times the series, ~1.1 MB per table at the
	// north-star size) live in HBM: written once here, read once by k_horner.
	template <int L>
	__global__ void __launch_bounds__(128, 4)
	    k_scale(DevCtx cx, int n_tables, const TableRef* __restrict__ tref, const mpf<L>* __restrict__ delta,
	            const mpf<L>* __restrict__ b, const mpf<L>* __restrict__ pw, mpf<L>* __restrict__ a)
	{
		const uint32_t n1 = cx.n_max + 1;
		const int64_t gid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
		const int t = int(gid / n1);
		const uint32_t n = uint32_t(gid % n1);
		if (t >= n_tables) return;
		const int nk = int(cx.lambda) + 1;
		const int prec = cx.prec;
		const TableRef tr = tref[t];
		mpf<L>* at = a + size_t(t) * nk * n1 + n;
		mpf<L> cur, q;
		{
			const mpf<L>& c4 = pw[size_t(t) * (nk + 1)];
			if (tr.job1 >= 0)
			{
				// divergent operator: average of the two shifted series               (hor_recursion.cpp:23-27)
				mpf<L> u, v;
				copy(u, b[size_t(tr.job0) * n1 + n]);
				copy(v, b[size_t(tr.job1) * n1 + n]);
				add(u, u, v, prec);
				mul_2si(u, u, -1);
				fast::mul(cur, u, c4, prec);
			}
			else
				fast::mul(cur, b[size_t(tr.job0) * n1 + n], c4, prec);
		}
		copy(at[0], cur);
		copy(q, delta[t]);
		for (int k = 1; k < nk; ++k)
		{
			mpf<L> f, nxt;
			fast::add_small(f, q, int64_t(n), prec);   // q_{k-1} + n
			fast::mul(nxt, cur, f, prec);
			copy(at[size_t(k) * n1], nxt);
			copy(cur, nxt);
			fast::add_small(q, q, -1, prec);           // pow_ -= 1 (real_function.hpp:239)
		}
	}

	// rho-derivative stage, part 2: Horner sums s_k = sum_n a_k[n] rho^n (approximate, real_function.hpp:248-259) and
	// f_k = s_k * rho^(q_k) / k! (real_function.hpp:262, hor_recursion.cpp:46-53).  One thread per (table, k).
	template <int L>
	__global__ void __launch_bounds__(128, 4)
	    k_horner(DevCtx cx, int n_tables, const mpf<L>* __restrict__ a, const mpf<L>* __restrict__ pw, mpf<L>* __restrict__ fr)
	{
		const int nk = int(cx.lambda) + 1;
		const int64_t gid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
		const int t = int(gid / nk), k = int(gid % nk);
		if (t >= n_tables) return;
		const uint32_t n1 = cx.n_max + 1;
		const int prec = cx.prec;
		const mpf<L>* ak = a + (size_t(t) * nk + k) * n1;
		mpf<L> rho, s;
		copy(rho, *reinterpret_cast<const mpf<L>*>(cx.rho));
		copy(s, ak[cx.n_max]);
		for (uint32_t n = cx.n_max; n-- > 0;) fast::fma(s, s, rho, ak[n], prec);
		mpf<L> rp, kf, f;
		copy(rp, pw[size_t(t) * (nk + 1) + 1 + k]);
		copy(kf, reinterpret_cast<const mpf<L>*>(cx.kfact)[k]);
		deriv_finish(f, s, rp, kf, uint32_t(k), prec);
		copy(fr[size_t(t) * nk + k], f);
	}

}  // namespace qb
