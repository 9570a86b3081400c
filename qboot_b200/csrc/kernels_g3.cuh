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

namespace qb
{
	// Stage a5 + a6 as a short sequence of WIDE kernels (every lane does the same amount of work):
	//   k_rho_to_z        thread per (table, column pair i / lambda - i): entries of the rho -> z matrix-vector product (dot
	//                     products of lambda + 1 terms, matrix.hpp:444-456); the i == 0 thread also leaves 4 C_2 of the table in qc[table]
	//   k_offdiag_val     per dy-level n, thread per (table, dx = m): the part of the Casimir recursion that only reads
	//                     levels n - 1 and n - 2 (context.cpp:72-121) -> written into the table slot (m, n)
	//   k_offdiag_close   per dy-level n, thread per table: the same-level correction, sequential in m (context.cpp:122-131)
	// lambda / 2 levels => 1 + 2 * (lambda / 2) launches per chunk; the kernel boundary is the level barrier.
	// (thread per (table, pair of columns): column i and column lambda - i of the triangular matrix together cost lambda + 2
	// products whatever i is, so the lanes of a warp finish together -- one thread per column left 22 of 32 lanes idle)
	template <int L>
	__global__ void __launch_bounds__(128) k_rho_to_z(DevCtx cx, int n_tables, const mpf<L>* __restrict__ delta,
	                                                  const uint32_t* __restrict__ spin, const mpf<L>* __restrict__ fr,
	                                                  mpf<L>* __restrict__ g, mpf<L>* __restrict__ qc)
	{
		const int lambda = int(cx.lambda), nk = lambda + 1, np = (nk + 1) / 2, dimM = cf_dim(lambda), prec = cx.prec;
		const int gid = blockIdx.x * blockDim.x + threadIdx.x;
		const int t = gid / np, i = gid - t * np;
		if (t >= n_tables) return;
		const mpf<L>* M = reinterpret_cast<const mpf<L>*>(cx.mat);
		QB_ROLL
		for (int w = 0; w < 2; ++w)
		{
			const int col = w == 0 ? i : lambda - i;
			if (w == 1 && col == i) break;
			mpf<L> z;
			rho_to_z_entry(z, fr + size_t(t) * nk, 1, M, lambda, col, prec);
			copy(g[size_t(t) * dimM + col], z);
		}
		if (i == 0)
		{
			mpf<L> lD, q4;
			copy(lD, delta[t]);
			quad_casimir4(q4, lD, spin[t], cx.eps, prec);
			copy(qc[t], q4);
		}
	}

#ifndef QB_OFFDIAG_T
#define QB_OFFDIAG_T 128
#endif
#ifndef QB_OFFDIAG_CTAS
#define QB_OFFDIAG_CTAS 4
#endif
	constexpr int QB_OFFDIAG_THREADS = QB_OFFDIAG_T;
	template <int L>
	__global__ void __launch_bounds__(QB_OFFDIAG_THREADS, QB_OFFDIAG_CTAS) k_offdiag_val(DevCtx cx, int n_tables, int n, const mpf<L>* __restrict__ S,
	                                                                    const mpf<L>* __restrict__ P, const mpf<L>* __restrict__ qc, mpf<L>* g)
	{
		const int lambda = int(cx.lambda), dimM = cf_dim(lambda), cnt = lambda - 2 * n + 1;
		const int gid = blockIdx.x * blockDim.x + threadIdx.x;
		int t = gid / cnt;
		const int m = gid - t * cnt;
		const bool active = t < n_tables;  // idle threads keep the barriers of the lockstep interpreter company
		if (!active) t = 0;
		mpf<L>* f = g + size_t(t) * dimM;
		mpf<L> lS, lP, q4, val;
		copy(lS, S[t]);
		copy(lP, P[t]);
		copy(q4, qc[t]);
		offdiag_val<L, true>(val, f, 1, lambda, m, n, lS, lP, q4, cx.eps, cx.prec, active);
		if (active) copy(f[cf_index(lambda, m, n)], val);
	}

	template <int L>
	__global__ void __launch_bounds__(128) k_offdiag_close(DevCtx cx, int n_tables, int n, mpf<L>* g)
	{
		const int lambda = int(cx.lambda), dimM = cf_dim(lambda), cnt = lambda - 2 * n + 1;
		const int t = blockIdx.x * blockDim.x + threadIdx.x;
		if (t >= n_tables) return;
		mpf<L>* f = g + size_t(t) * dimM + cf_index(lambda, 0, n);
		mpf<L> p1, p2, p3, cur, val;
		set_zero(p1);
		set_zero(p2);
		set_zero(p3);
		QB_ROLL
		for (int m = 0; m < cnt; ++m)
		{
			copy(val, f[m]);
			offdiag_close(cur, m, val, p1, p2, p3, cx.prec);
			copy(f[m], cur);
			copy(p3, p2);
			copy(p2, p1);
			copy(p1, cur);
		}
	}

}  // namespace qb
