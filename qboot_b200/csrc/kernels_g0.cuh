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
		mpf<L>* out = coef + size_t(job) * QB_NCOEF;
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
			copy(out[c], r);
			return;
		}
		// p_0 and p_{K-1}: products of linear factors; the d == 0 thread of each builds the whole polynomial
		if (d != 0) return;
		mpf<L> pe[5];
		rec_coeffs_end(pe, lD, sp, cx.eps, cx.prec, i == 0);
		for (int k = 0; k < st; ++k) copy(out[i * st + k], pe[k]);
	}

	// 64-bit record transfers (see kernels_g2.cuh): for even L every record of a 256-byte aligned array is 8-byte aligned
	template <int L>
	__device__ __forceinline__ void load_rec64(mpf<L>& d, const mpf<L>* src)
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
	template <int L>
	__device__ __forceinline__ void store_rec64(mpf<L>* dst, const mpf<L>& s)
	{
		if constexpr (L % 2 == 0)
		{
			uint2* dp = reinterpret_cast<uint2*>(dst);
			dp[0] = make_uint2(s.sk, uint32_t(s.exp));
#pragma unroll
			for (int j = 0; j < L / 2; ++j) dp[1 + j] = make_uint2(s.m[2 * j], s.m[2 * j + 1]);
		}
		else
			copy(*dst, s);
	}

	// values of the K recurrence polynomials at every n: pv[(job * n_max + n - 1) * 8 + i] = p_i(n), by Horner with the
	// integer n (include/qboot/algebra/polynomial.hpp:66-83).  They do not depend on the series, so they are taken off its
	// sequential chain: one thread per (job, n).
	template <int L>
	__global__ void k_polyvals(DevCtx cx, int n_jobs, const uint32_t* __restrict__ spin, const mpf<L>* __restrict__ coef,
	                           mpf<L>* __restrict__ pv)
	{
		const int64_t gid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
		const int job = int(gid / cx.n_max);
		const uint32_t n = uint32_t(gid % cx.n_max) + 1u;
		if (job >= n_jobs) return;
		const uint32_t sp = spin[job];
		int K = sp == 0 ? 6 : 8, deg = sp == 0 ? 3 : 4;
#ifdef __CUDA_ARCH__
		// keep ONE copy of the loop body: without this the compiler clones and unrolls it per spin class and the kernel
		// outgrows the instruction cache (measured: 75 % of its stall samples were instruction fetch)
		asm volatile("" : "+r"(K), "+r"(deg));
#endif
		const int st = deg + 1;
		const mpf<L>* c = coef + size_t(job) * QB_NCOEF;
		mpf<L>* out = pv + size_t(gid) * 8;
		QB_ROLL
		for (int i = 0; i < K; ++i)
		{
			mpf<L> s, cd;
			load_rec64(s, &c[i * st + deg]);
			QB_ROLL
			for (int d = deg - 1; d >= 0; --d)
			{
				fast::mul_small(s, s, int64_t(n), cx.prec);
				load_rec64(cd, &c[i * st + d]);
				fast::add(s, s, cd, cx.prec);
			}
			store_rec64(&out[i], s);
		}
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
			pow_from_log<L>(r, *reinterpret_cast<const mpf<E>*>(cx.lnrhoe), y, *reinterpret_cast<const mpf<E>*>(cx.ln2e), prec);
			break;
		case 14:  // 4^x
			pow_from_log<L>(r, *reinterpret_cast<const mpf<E>*>(cx.ln4e), x, *reinterpret_cast<const mpf<E>*>(cx.ln2e), prec);
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
