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

	// Powers of one table: pw[t][0] = 4^Delta (src/hor_recursion.cpp:43), pw[t][1 + k] = rho^(q_k) with q_k = Delta - k by
	// k rounded decrements (include/qboot/algebra/real_function.hpp:239,262), and the q_k themselves in
	// pw[t][lambda + 2 + k] (k_scale multiplies by q_{k-1} + n).  Row length 2 (lambda + 1) + 1.  One thread per table: two exponentials
	// in the extended format (4^Delta and X = rho^(q_0)), then for every k
	//     rho^(q_k) = X * rho^-k * (1 + eta_k ln rho),   eta_k = q_k - (q_0 - k)   (|eta_k| <~ k 2^-prec |q|, eta^2 negligible)
	// with rho^-k precomputed in the extended format (DevCtx::rhoinv): 3 extended products per power instead of an exp.
	template <int L>
	__global__ void k_pow(DevCtx cx, int n_tables, const mpf<L>* __restrict__ delta, mpf<L>* __restrict__ pw)
	{
		constexpr int E = L + QB_EXT_LIMBS;
		const int pe = 32 * E;
		const int np = int(cx.lambda) + 2;
		const int row = 2 * int(cx.lambda) + 3;
		int t = blockIdx.x * blockDim.x + threadIdx.x;
		if (t >= n_tables) return;
		mpf<L> q0, q, r;
		copy(q0, delta[t]);
		copy(pw[size_t(t) * row + np], q0);
		pow_ziv<L>(r, cx.ln4e, cx.ln2e, cx.ln4w, cx.ln2w, q0, cx.prec, cx.ziv_slack, cx.ziv);
		copy(pw[size_t(t) * row], r);
		// X = exp(q0 ln rho) in the extended format
		mpf<E> X, lnr, l2, qe, te, corr, one, v;
		copy(lnr, *reinterpret_cast<const mpf<E>*>(cx.lnrhoe));
		copy(l2, *reinterpret_cast<const mpf<E>*>(cx.ln2e));
		widen(qe, q0);
		fast::mul(te, qe, lnr, pe);
		exp_ext(X, te, l2);
		narrow(r, X, cx.prec);
		// every power is rounded from an extended value: a rounding in doubt (ziv_safe) is redone from a full
		// exponential at the second-level width
		auto redo = [&](mpf<L>& out, const mpf<L>& y) {
			constexpr int E2 = L + QB_EXT2_LIMBS;
			ziv_count(cx.ziv, 0);
			if (!pow_from_log<L, QB_EXT2_LIMBS>(out, *reinterpret_cast<const mpf<E2>*>(cx.lnrhow), y, *reinterpret_cast<const mpf<E2>*>(cx.ln2w),
			                                    cx.prec))
				ziv_count(cx.ziv, 1);
		};
		if (!ziv_safe(X, cx.prec, cx.ziv_slack)) redo(r, q0);
		copy(pw[size_t(t) * row + 1], r);
		set_si(one, 1, pe);
		copy(q, q0);
		const mpf<E>* rinv = reinterpret_cast<const mpf<E>*>(cx.rhoinv);
		for (int k = 1; k + 1 < np; ++k)
		{
			add_si(q, q, -1, cx.prec);                 // q_k, rounded as the reference rounds it
			copy(pw[size_t(t) * row + np + k], q);
			// eta = q_k - q_0 + k, exact in the extended format
			mpf<E> qk, eta;
			widen(qk, q);
			sub(eta, qk, qe, pe);
			add_si(eta, eta, k, pe);
			fast::mul(v, X, rinv[k], pe);
			if (!is_zero(eta))
			{
				mul(corr, eta, lnr, pe);
				add(corr, corr, one, pe);
				fast::mul(v, v, corr, pe);
			}
			narrow(r, v, cx.prec);
			if (!ziv_safe(v, cx.prec, cx.ziv_slack + 2)) redo(r, q);  // three more roundings and the dropped eta^2 term: 2 bits wider
			copy(pw[size_t(t) * row + 1 + k], r);
		}
	}

	template <int L>
	__global__ void k_vtod(DevCtx cx, int n_d, const mpf<L>* __restrict__ d, mpf<L>* __restrict__ v)
	{
		int i = blockIdx.x * blockDim.x + threadIdx.x;
		if (i >= n_d) return;
		constexpr int E = L + QB_EXT_LIMBS;
		mpf<L> md, p4;
		copy(md, d[i]);
		neg(md);
		pow_ziv<L>(p4, cx.ln4e, cx.ln2e, cx.ln4w, cx.ln2w, md, cx.prec, cx.ziv_slack, cx.ziv);
		v_to_d_table(v + size_t(i) * cf_dim(int(cx.lambda)), 1, int(cx.lambda), d[i], p4, cx.prec);
	}

	template <int L>
	__global__ void k_fblock(DevCtx cx, int n_terms, int dim_max, const TermRef* __restrict__ terms,
	                         const mpf<L>* __restrict__ g, const mpf<L>* __restrict__ v, mpf<L>* __restrict__ out)
	{
		int gid = blockIdx.x * blockDim.x + threadIdx.x;
		int term = gid / dim_max, e = gid % dim_max;
		if (term >= n_terms) return;
		TermRef tr = terms[term];
		const int lambda = int(cx.lambda), dimM = cf_dim(lambda);
		if (e >= cf_dim_sym(lambda, tr.sym)) return;
		// e-th kept entry in (N-major, M-minor) order with M = sym mod 2
		int N = 0, rem = e;
		for (;; ++N)
		{
			int top = lambda - 2 * N - tr.sym;
			int row = top < 0 ? 0 : top / 2 + 1;  // number of M in this row with the right parity
			if (rem < row) break;
			rem -= row;
		}
		int M = 2 * rem + tr.sym;
		mpf<L> r;
		fblock_entry(r, v + size_t(tr.dix) * dimM, 1, g + size_t(tr.table) * dimM, 1, lambda, M, N, cx.prec);
		copy(out[size_t(tr.out_ofs) + e], r);
	}

}  // namespace qb
