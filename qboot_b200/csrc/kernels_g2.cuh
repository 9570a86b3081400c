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
	// Both kernels of this stage keep the running value of their chain -- the row operand of every product -- in SHARED
	// memory, limb-major (limb j of thread x at sh[j * 128 + x], conflict-free), with its sign and exponent in registers:
	// the rolled product loop fetches two limbs per iteration, and a thread-local record would live in local memory,
	// whose footprint (resident threads x stack) overflows L1 and spills to DRAM (measured: 7 GB per launch).
	constexpr int QB_DERIV_THREADS = 128;

	template <int L>
	__device__ __forceinline__ void chain_store(uint32_t* my, uint32_t& sk, int32_t& exp, const mpf<L>& v)
	{
		sk = v.sk;
		exp = v.exp;
#pragma unroll
		for (int j = 0; j < L; ++j) my[j * QB_DERIV_THREADS] = v.m[j];
	}
	template <int L>
	__device__ __forceinline__ void chain_load(mpf<L>& v, const uint32_t* my, uint32_t sk, int32_t exp)
	{
		v.sk = sk;
		v.exp = exp;
#pragma unroll
		for (int j = 0; j < L; ++j) v.m[j] = my[j * QB_DERIV_THREADS];
	}

	// 64-bit record transfers: records are 4 (L + 2) bytes apart in arrays whose base is 256-byte aligned, so for even L
	// every record is 8-byte aligned and moves as (L + 2) / 2 double words -- half the load / store instructions of the
	// word-wise copy() (k_scale was throttled by its memory-instruction queue: 34 STG + 34 LDG + 32 STS + 32 LDS per step).
	template <int L>
	__device__ __forceinline__ void load_rec(mpf<L>& d, const mpf<L>* src)
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
	__device__ __forceinline__ void store_rec(mpf<L>* dst, const mpf<L>& s)
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

	// rho-derivative stage, part 1: every scaled coefficient array a_k[n], k = 0..lambda (reference: h *= 4^Delta,
	// src/hor_recursion.cpp:43, then lambda times RealFunctionWithPower::derivate, real_function.hpp:236-240):
	//   a_0[n] = b[n] * 4^Delta,   a_k[n] = a_{k-1}[n] * (q_{k-1} + n),   q_k = q_{k-1} - 1,  q_0 = Delta  (q_k from k_pow)
	// One thread per (table, n) walks k: no dependency between threads, no barrier.  Layout a[table][k][n] (n fastest) so
	// that part 2 streams each chain contiguously.  The arrays (lambda+1 times the series, ~1.1 MB per table at the
	// north-star size) live in HBM: written once here, read once by k_horner.
	template <int L>
	__global__ void __launch_bounds__(QB_DERIV_THREADS, 4)
	    k_scale(DevCtx cx, int n_tables, const TableRef* __restrict__ tref, const mpf<L>* __restrict__ b,
	            const mpf<L>* __restrict__ pw, mpf<L>* __restrict__ a)
	{
		__shared__ uint32_t sh[L * QB_DERIV_THREADS];
		uint32_t* my = sh + threadIdx.x;
		const uint32_t n1 = cx.n_max + 1;
		const int64_t gid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
		const int t = int(gid / n1);
		const uint32_t n = uint32_t(gid % n1);
		if (t >= n_tables) return;
		const int nk = int(cx.lambda) + 1;
		const int prec = cx.prec;
		const TableRef tr = tref[t];
		mpf<L>* at = a + size_t(t) * nk * n1 + n;
		const mpf<L>* pt = pw + size_t(t) * (2 * nk + 1);  // [0] 4^Delta, [1 + k] rho^(q_k), [nk + 1 + k] q_k
		uint32_t csk;
		int32_t cexp;
		{
			mpf<L> cur;
			if (tr.job1 >= 0)
			{
				// divergent operator: average of the two shifted series               (hor_recursion.cpp:23-27)
				mpf<L> u, v;
				copy(u, b[size_t(tr.job0) * n1 + n]);
				copy(v, b[size_t(tr.job1) * n1 + n]);
				fast::add(u, u, v, prec);
				mul_2si(u, u, -1);
				fast::mul(cur, u, pt[0], prec);
			}
			else
				fast::mul(cur, b[size_t(tr.job0) * n1 + n], pt[0], prec);
			store_rec(&at[0], cur);
			chain_store(my, csk, cexp, cur);
		}
		QB_ROLL
		for (int k = 1; k < nk; ++k)
		{
			mpf<L> f, nxt, qk;
			load_rec(qk, &pt[nk + k]);
			fast::add_small(f, qk, int64_t(n), prec);   // q_{k-1} + n
			if ((csk & 3u) == K_REG && (f.sk & 3u) == K_REG)
				fast::mul_strided<L, QB_DERIV_THREADS>(nxt, csk, cexp, my, f, prec);
			else
			{
				mpf<L> cur;
				chain_load(cur, my, csk, cexp);
				fast::cold_mul(nxt, cur, f, prec);
			}
			store_rec(&at[size_t(k) * n1], nxt);
			chain_store(my, csk, cexp, nxt);
		}
	}

	// rho-derivative stage, part 2: Horner sums s_k = sum_n a_k[n] rho^n (approximate, real_function.hpp:248-259) and
	// f_k = s_k * rho^(q_k) / k! (real_function.hpp:262, hor_recursion.cpp:46-53).  One thread per (table, k).
	template <int L>
	__global__ void __launch_bounds__(QB_DERIV_THREADS, 4)
	    k_horner(DevCtx cx, int n_tables, const mpf<L>* __restrict__ a, const mpf<L>* __restrict__ pw, mpf<L>* __restrict__ fr)
	{
		__shared__ uint32_t sh[L * QB_DERIV_THREADS];
		uint32_t* my = sh + threadIdx.x;
		const int nk = int(cx.lambda) + 1;
		const int64_t gid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
		const int t = int(gid / nk), k = int(gid % nk);
		if (t >= n_tables) return;
		const uint32_t n1 = cx.n_max + 1;
		const int prec = cx.prec;
		const mpf<L>* ak = a + (size_t(t) * nk + k) * n1;
		mpf<L> rho;
		copy(rho, *reinterpret_cast<const mpf<L>*>(cx.rho));
		uint32_t ssk;
		int32_t sexp;
		{
			mpf<L> s;
			copy(s, ak[cx.n_max]);
			chain_store(my, ssk, sexp, s);
		}
		QB_ROLL
		for (uint32_t n = cx.n_max; n-- > 0;)
		{
			mpf<L> r;
			const mpf<L>& c = ak[n];
			bool done = false;
			if ((ssk & 3u) == K_REG && (c.sk & 3u) == K_REG) done = fast::fma_strided<L, QB_DERIV_THREADS>(r, ssk, sexp, my, rho, c, prec);
			if (!done)
			{
				mpf<L> s;
				chain_load(s, my, ssk, sexp);
				fast::cold_fma(r, s, rho, c, prec);
			}
			chain_store(my, ssk, sexp, r);
		}
		mpf<L> s, rp, kf, f;
		chain_load(s, my, ssk, sexp);
		copy(rp, pw[size_t(t) * (2 * nk + 1) + 1 + k]);
		copy(kf, reinterpret_cast<const mpf<L>*>(cx.kfact)[k]);
		deriv_finish(f, s, rp, kf, uint32_t(k), prec);
		copy(fr[size_t(t) * nk + k], f);
	}

}  // namespace qb
