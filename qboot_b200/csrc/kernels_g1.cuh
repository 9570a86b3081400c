// qboot_b200 -- CUDA kernels of the block-table pipeline (sm_100a), templated on the limb count L; split by kernel
// group so that each translation unit (kernel_inst.cu -DQB_KGROUP=g) only parses what it instantiates:
//   kernels_g0.cuh  k_rec_coeffs  one thread per (series job, recurrence coefficient)              -> coef[job][40]
//                   k_ops         diagnostic: one primitive elementwise
//   kernels_g1.cuh  k_series      32 series jobs per CTA: a producer warp (recurrence polynomials at n+1) feeding a consumer warp (the chain) -> b[job][0..n_max]
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
	// record <-> shared memory, word-major with the 32 lanes of a warp side by side (conflict free): word w of lane l at base[w * 32 + l]
	template <int L>
	__device__ __forceinline__ void sts_rec(uint32_t* base, const mpf<L>& s)
	{
		base[0] = s.sk;
		base[32] = uint32_t(s.exp);
#pragma unroll
		for (int j = 0; j < L; ++j) base[(2 + j) * 32] = s.m[j];
	}
	template <int L>
	__device__ __forceinline__ void lds_rec(mpf<L>& d, const uint32_t* base)
	{
		d.sk = base[0];
		d.exp = int32_t(base[32]);
#pragma unroll
		for (int j = 0; j < L; ++j) d.m[j] = base[(2 + j) * 32];
	}
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

	// The series recurrence b[n] = -(sum_{i=1}^{min(n,K-1)} p_i(n) b[n-i]) / p_0(n)       (src/hor_recursion.cpp:31-37)
	// for 32 jobs per CTA, as a two-warp producer / consumer pair:
	//   warp 1 (producer) evaluates the K recurrence polynomials at n + 1 by Horner with the integer n, exactly as the
	//          reference's Polynomial::eval does (include/qboot/algebra/polynomial.hpp:66-83: mul_ui, add, each rounded),
	//          into one half of a double buffer in shared memory;
	//   warp 0 (consumer) runs the dependent chain of step n -- K - 1 fused multiply-adds and one division -- reading
	//          the p_i(n) of the other half.
	// The polynomial values are independent of the series, so the producer is never on the critical path, and they never
	// travel through HBM (a separate k_polyvals kernel used to write 435 KB per table and cost 12 % of the kernel time).
	constexpr int QB_SERIES_THREADS = 64;
	template <int L>
	constexpr size_t series_smem_bytes()
	{
		return size_t(2) * 8 * (L + 2) * 32 * sizeof(uint32_t);
	}
	template <int L>
	__global__ void __launch_bounds__(QB_SERIES_THREADS)
	    k_series(DevCtx cx, int n_jobs, const mpf<L>* __restrict__ delta, const uint32_t* __restrict__ spin, const mpf<L>* __restrict__ b0,
	             const mpf<L>* __restrict__ coef, mpf<L>* __restrict__ b)
	{
		extern __shared__ uint32_t pvbuf[];
		constexpr int W = L + 2;
		const int lane = threadIdx.x & 31;
		const bool producer = (threadIdx.x >> 5) != 0;
		const int job = blockIdx.x * 32 + lane;
		const int prec = cx.prec;
		const uint32_t n_max = cx.n_max;
		bool active = job < n_jobs;
		mpf<L>* bj = b + size_t(active ? job : 0) * (n_max + 1);
		if (active && is_zero(delta[job]))
		{
			// unit operator: b = (b0, 0, 0, ...)                                  (src/hor_recursion.cpp:20)
			if (!producer)
			{
				copy(bj[0], b0[job]);
				for (uint32_t n = 1; n <= n_max; ++n) set_zero(bj[n]);
			}
			active = false;
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
		const mpf<L>* c = coef + size_t(active ? job : 0) * QB_NCOEF;
		auto slot = [&](uint32_t n, int i) { return pvbuf + ((size_t(n & 1u) * 8 + size_t(i)) * W) * 32 + lane; };
		auto produce = [&](uint32_t n) {
			QB_ROLL
			for (int i = 0; i < K; ++i)
			{
				mpf<L> s, cd;
				ldg_rec64(s, &c[i * st + deg]);
				QB_ROLL
				for (int d = deg - 1; d >= 0; --d)
				{
					fast::mul_small(s, s, int64_t(n), prec);
					ldg_rec64(cd, &c[i * st + d]);
					fast::add(s, s, cd, prec);
				}
				sts_rec(slot(n, i), s);
			}
		};
		if (active)
		{
			if (producer)
				produce(1u);
			else
				copy(bj[0], b0[job]);
		}
		__syncthreads();
		for (uint32_t n = 1; n <= n_max; ++n)
		{
			if (active)
			{
				if (producer)
				{
					if (n < n_max) produce(n + 1u);
				}
				else
				{
					mpf<L> sum, bn, p, bp;
					lds_rec(p, slot(n, 1));
					ldg_rec64(bp, &bj[n - 1]);
					fast::mul(sum, p, bp, prec);  // fma(0, p_1 b[n-1]) rounds like the product
					QB_ROLL
					for (int i = 2; i < K; ++i)
						if (uint32_t(i) <= n)
						{
							lds_rec(p, slot(n, i));
							ldg_rec64(bp, &bj[n - uint32_t(i)]);
							fast::fma(sum, p, bp, sum, prec);
						}
					neg(sum);
					lds_rec(p, slot(n, 0));
					fast::div(bn, sum, p, prec);
					copy(bj[n], bn);
				}
			}
			__syncthreads();
		}
	}

}  // namespace qb
