// qboot_b200 -- CUDA kernels of the PMP stage (sm_100a), templated on the limb count L (see kernels_g4.cuh for the
// conventions of the block-table kernels):
//   k_pmp_accum  one thread per (constraint block, matrix position rc, variable n, sample k), k fastest: the 32 lanes of
//                a warp evaluate the SAME entry of the same terms at 32 different sample points, so they walk identical
//                loops (F-block dot products of equal length) over different block tables         -> M[block][p][n]
//   k_pmp_pack   one thread per (block, row p, free variable m | the d_c column), m fastest         -> d_B, d_c
//   k_poly_eval  one thread per (polynomial, point, scale): fma Horner, then the scale                -> out
//   k_dec_format one thread per number: decimal digits + layout into a fixed-size slot              -> slots, lengths
//   k_dec_gather one warp per number: slot -> its final place in the contiguous text               -> text
#pragma once

#include <cuda_runtime.h>

#define QB_TABLE_QUAL static __device__
#include "launch.cuh"
#include "pmp.cuh"

namespace qb
{
	// largest b with first[b] <= gid, where first(b) is `thread0` or `pack0` of the block descriptors
	template <bool PACK>
	__device__ __forceinline__ int pmp_find_block(const PmpBlock* __restrict__ blocks, int n_blocks, int64_t gid)
	{
		int lo = 0, hi = n_blocks - 1;
		while (lo < hi)
		{
			int mid = (lo + hi + 1) >> 1;
			int64_t f = PACK ? blocks[mid].pack0 : blocks[mid].thread0;
			if (f <= gid)
				lo = mid;
			else
				hi = mid - 1;
		}
		return lo;
	}

	template <int L>
	__global__ void k_pmp_accum(DevCtx cx, int n_blocks, const PmpBlock* __restrict__ blocks, const PmpTerm* __restrict__ terms,
	                            const int32_t* __restrict__ tab, int n_eq, const int32_t* __restrict__ eq_ofs,
	                            const int32_t* __restrict__ eq_sym, const mpf<L>* __restrict__ g, const mpf<L>* __restrict__ v,
	                            const mpf<L>* __restrict__ coef, mpf<L>* __restrict__ Marena, int64_t total)
	{
		const int64_t gid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
		if (gid >= total) return;
		const PmpBlock blk = blocks[pmp_find_block<false>(blocks, n_blocks, gid)];
		const int N = eq_ofs[n_eq], ns = blk.n_samples;
		int64_t t = gid - blk.thread0;
		const int k = int(t % ns);
		t /= ns;
		const int n = int(t % N), rc = int(t / N);
		int eq = 0;
		while (eq + 1 < n_eq && eq_ofs[eq + 1] <= n) ++eq;
		mpf<L> r;
		pmp_entry<L>(r, terms + blk.term0, blk.n_terms, eq, rc, k, eq_sym[eq], n - eq_ofs[eq], tab, g, v, coef, int(cx.lambda), cx.prec);
		copy(Marena[blk.ofs + (int64_t(rc) * ns + k) * N + n], r);
	}

	template <int L>
	__global__ void k_pmp_pack(DevCtx cx, int n_blocks, const PmpBlock* __restrict__ blocks, int N, int n_free,
	                           const int32_t* __restrict__ free_idx, int n_elim, const int32_t* __restrict__ lead,
	                           const mpf<L>* __restrict__ eqn, const mpf<L>* __restrict__ eqn_target, const mpf<L>* __restrict__ Marena,
	                           const mpf<L>* __restrict__ targets, mpf<L>* __restrict__ Barena, mpf<L>* __restrict__ Carena, int64_t total)
	{
		const int64_t gid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
		if (gid >= total) return;
		const PmpBlock blk = blocks[pmp_find_block<true>(blocks, n_blocks, gid)];
		int64_t t = gid - blk.pack0;
		const int m = int(t % (n_free + 1));
		const int64_t p = t / (n_free + 1);
		mpf<L> r;
		pmp_pack_entry<L>(r, Marena + blk.ofs + p * N, m, n_free, free_idx, n_elim, lead, eqn, eqn_target,
		                  blk.ofs_t >= 0 ? targets + blk.ofs_t + p : nullptr, N, cx.prec);
		if (m < n_free)
			copy(Barena[blk.ofs_b + p * n_free + m], r);
		else
			copy(Carena[blk.ofs_c + p], r);
	}

	// polynomials at points, scaled (bilinear bases at the sample points): one thread per item
	template <int L>
	__global__ void k_poly_eval(DevCtx cx, int n_items, const EvalItem* __restrict__ items, const int32_t* __restrict__ poly_ofs,
	                            const mpf<L>* __restrict__ coef, const mpf<L>* __restrict__ x, const mpf<L>* __restrict__ s,
	                            mpf<L>* __restrict__ out)
	{
		const int i = blockIdx.x * blockDim.x + threadIdx.x;
		if (i >= n_items) return;
		const EvalItem it = items[i];
		const int o = poly_ofs[it.poly], deg = poly_ofs[it.poly + 1] - o - 1;
		mpf<L> r;
		poly_eval_scaled<L>(r, coef + o, deg, x[it.x], s[it.s], cx.prec);
		copy(out[i], r);
	}

	// numbers -> text slots of `slot` bytes each; len[i] = bytes used, or < 0 when the host has to format number i
	template <int L>
	__global__ void k_dec_format(DecTables tb, int64_t n, const mpf<L>* __restrict__ x, char* __restrict__ slots, int slot,
	                             int32_t* __restrict__ len)
	{
		const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
		if (i >= n) return;
		mpf<L> v;
		copy(v, x[i]);
		len[i] = format_decimal<L>(slots + i * slot, v, tb);
	}
	// text[ofs[i] .. ofs[i] + len[i]) = slot i; one warp per number
	static __global__ void k_dec_gather(int64_t n, const char* __restrict__ slots, int slot, const int32_t* __restrict__ len,
	                                    const int64_t* __restrict__ ofs, char* __restrict__ text)
	{
		const int64_t w = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
		const int lane = threadIdx.x & 31;
		if (w >= n) return;
		const int l = len[w];
		const char* src = slots + w * slot;
		char* dst = text + ofs[w];
		for (int j = lane; j < l; j += 32) dst[j] = src[j];
	}
}  // namespace qb
