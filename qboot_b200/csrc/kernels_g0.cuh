// qboot_b200 -- CUDA kernels of the block-table pipeline (sm_100a), templated on the limb count L; split by kernel
// group so that each translation unit (kernel_inst.cu -DQB_KGROUP=g) only parses what it instantiates:
//   kernels_g0.cuh  k_rec_coeffs  one thread per (series job, recurrence coefficient)              -> coef, word-planar across jobs
//                   k_ops         diagnostic: one primitive elementwise
//   kernels_g1.cuh  k_series      one thread per series job, n_max sequential steps                -> b[job][0..n_max]
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
#include "hor_table.inc"

namespace qb
{
	// recurrence coefficients a[i][d] of one job (_get_rec_coeffs, include/qboot/hor_formula.hpp:34-39): one thread per
	// (coefficient, job) with the JOB as the fast index, so that the lanes of a warp walk the same monomial list and stay
	// converged; the d == 0 threads of p_0 and p_{K-1} build those products of linear factors
	template <int L>
	__global__ void k_rec_coeffs(DevCtx cx, int n_jobs, const mpf<L>* __restrict__ delta, const uint32_t* __restrict__ spin,
	                             const mpf<L>* __restrict__ S, const mpf<L>* __restrict__ P, mpf<L>* __restrict__ coef)
	{
		int gid = blockIdx.x * blockDim.x + threadIdx.x;
		int c = gid / n_jobs, job = gid - c * n_jobs;
		if (c >= QB_NCOEF) return;
		uint32_t sp = spin[job];
		const int K = sp == 0 ? 6 : 8, deg = sp == 0 ? 3 : 4, st = deg + 1;
		if (c >= K * st) return;
		int i = c / st, d = c % st;
		uint32_t* out = reinterpret_cast<uint32_t*>(coef);
		const HorPoly* polys = sp == 0 ? QB_HOR_POLY_0 : QB_HOR_POLY_1;
		const HorMono* monos = sp == 0 ? QB_HOR_MONO_0 : QB_HOR_MONO_1;
		mpf<L> lD, lS, lP;
		copy(lD, delta[job]);
		if (i >= 1 && i < K - 1)
		{
			mpf<L> r;
			copy(lS, S[job]);
			copy(lP, P[job]);
			hor_coefficient(r, polys[(i - 1) * st + d], monos, lD, lS, lP, sp, cx.eps, cx.prec);
			coef_store(out, c, job, n_jobs, r);
			return;
		}
		// p_0 and p_{K-1}: products of linear factors; the d == 0 thread of each builds the whole polynomial
		if (d != 0) return;
		mpf<L> pe[5];
		rec_coeffs_end(pe, lD, sp, cx.eps, cx.prec, i == 0);
		for (int k = 0; k < st; ++k) coef_store(out, i * st + k, job, n_jobs, pe[k]);
	}

	// Elementwise primitive (diagnostic entry qb_op_batch): lets the GPU test-suite pin the device build of every
	// arithmetic routine -- generic (0..24) and register-resident fast paths (40..46) -- against MPFR.
	template <int L>
	__global__ void k_ops(DevCtx cx, int op, int n, const mpf<L>* __restrict__ a, const mpf<L>* __restrict__ b,
	                      const mpf<L>* __restrict__ c, const int64_t* __restrict__ ia, const int64_t* __restrict__ ib,
	                      mpf<L>* __restrict__ out)
	{
		int i = blockIdx.x * blockDim.x + threadIdx.x;
		if (i >= n) return;
		constexpr int E = L + QB_EXT_LIMBS;
		const int prec = cx.prec;
		mpf<L> x, y, z, r;
		set_zero(x);
		set_zero(y);
		set_zero(z);
		set_zero(r);
		if (a) copy(x, a[i]);
		if (b) copy(y, b[i]);
		if (c) copy(z, c[i]);
		int64_t p = ia ? ia[i] : 0, q = ib ? ib[i] : 1;
		switch (op)
		{
		case 0: add(r, x, y, prec); break;
		case 1: sub(r, x, y, prec); break;
		case 2: mul(r, x, y, prec); break;
		case 3: div(r, x, y, prec); break;
		case 4: fma(r, x, y, z, prec); break;
		case 6: mul_si(r, x, p, prec); break;
		case 7: add_si(r, x, p, prec); break;
		case 8: div_si(r, x, p, prec); break;
		case 9: mul_q(r, x, p, uint32_t(q), prec); break;
		case 10: div_q(r, x, p, uint32_t(q), prec); break;
		case 11: add_q(r, x, p, uint32_t(q), prec); break;
		case 12: set_q(r, p, uint32_t(q), prec); break;
		case 13:  // rho^y
			pow_ziv<L>(r, cx.lnrhoe, cx.ln2e, cx.lnrhow, cx.ln2w, y, prec, cx.ziv_slack, cx.ziv);
			break;
		case 14:  // 4^x
			pow_ziv<L>(r, cx.ln4e, cx.ln2e, cx.ln4w, cx.ln2w, x, prec, cx.ziv_slack, cx.ziv);
			break;
		case 18: si_sub(r, p, x, prec); break;
		case 20: sub_q(r, x, p, uint32_t(q), prec); break;
		case 40: fast::mul(r, x, y, prec); break;
		case 41: fast::fma(r, x, y, z, prec); break;
		case 42: fast::add(r, x, y, prec); break;
		case 43: fast::add(r, x, y, prec, true); break;
		case 44: fast::mul_small(r, x, p, prec); break;
		case 45: fast::add_small(r, x, p, prec); break;
		case 46: fast::div(r, x, y, prec); break;
		case 23: r.sk = uint32_t(cmp(x, y) + 1); break;
		case 24: r.sk = is_integer(x) ? 1u : 0u; break;
		default: set_nan(r); break;
		}
		copy(out[i], r);
	}

}  // namespace qb
