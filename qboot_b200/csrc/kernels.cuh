// qboot_b200 -- CUDA kernels of the block-table pipeline (sm_100a), templated on the limb count L.
//
//   k_rec_coeffs   one thread per (series job, recurrence coefficient)        -> coef[job][40]
//   k_series       one thread per series job, n_max sequential steps          -> b[job][0..n_max]
//   k_deriv        one CTA per G tables, (lambda+1) Horner chains per table running as a software pipeline through
//                  shared memory (chain k consumes the coefficient chain k-1 scaled one step earlier) -> fr[table][k]
//   k_finish       one thread per table: rho->z matrix and the transverse (Casimir) recursion        -> g[table][dim_M]
//   k_vtod         one thread per distinct d: 4^-d and the v^d table
//   k_fblock       one thread per (term, kept entry): 2 * proj(v^d * g)
//   k_ops          diagnostic: one primitive elementwise
//
// Numbers are records of L + 2 words (mpf<L>), arrays of records in HBM.  Arithmetic: mpf.cuh / hor.cuh / block.cuh.
// Staging rule: kernels copy operands from global / shared memory into thread-local records before any arithmetic
// call and copy results back afterwards (see hor.cuh).
#pragma once

#include <cuda_runtime.h>

#define QB_TABLE_QUAL static __device__
#include "block.cuh"
#include "hor_table.inc"

namespace qb
{
	// device-resident context constants (see qb_context_init)
	struct DevCtx
	{
		int prec;
		uint32_t n_max, lambda;
		Rational eps;
		const uint32_t* rho;      // mpf<L>
		const uint32_t* mat;      // (lambda+1)^2 mpf<L>, row-major
		const uint32_t* ln2e;     // mpf<L + QB_EXT_LIMBS>
		const uint32_t* ln4e;
		const uint32_t* lnrhoe;
		const uint32_t* kfact;    // k!, k = 0..lambda, mpf<L>
	};

	// a series job: one run of the recurrence at (Delta', spin, S, P)
	// a table: references one job, or two for a divergent operator (averaged, src/hor_recursion.cpp:21-28)
	struct TableRef
	{
		int32_t job0, job1;  // job1 < 0: single
	};

	constexpr int QB_NCOEF = 40;  // 8 polynomials x 5 coefficients (spin 0 uses 6 x 4 of them)

	template <int L>
	__global__ void k_rec_coeffs(DevCtx cx, int n_jobs, const mpf<L>* __restrict__ delta, const uint32_t* __restrict__ spin,
	                             const mpf<L>* __restrict__ S, const mpf<L>* __restrict__ P, mpf<L>* __restrict__ coef)
	{
		int gid = blockIdx.x * blockDim.x + threadIdx.x;
		int job = gid / QB_NCOEF, c = gid % QB_NCOEF;
		if (job >= n_jobs) return;
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

	template <int L>
	__global__ void k_series(DevCtx cx, int n_jobs, const mpf<L>* __restrict__ delta, const uint32_t* __restrict__ spin,
	                         const mpf<L>* __restrict__ b0, const mpf<L>* __restrict__ coef, mpf<L>* __restrict__ b)
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
		series<L>(bj, 1, cx.n_max, coef + size_t(job) * QB_NCOEF, spin[job], cx.prec);
	}

	// shared-memory pipeline of the lambda+1 derivative chains
	template <int L>
	__global__ void k_deriv(DevCtx cx, int n_tables, int G, const TableRef* __restrict__ tref,
	                        const mpf<L>* __restrict__ delta, const mpf<L>* __restrict__ b, mpf<L>* __restrict__ fr)
	{
		extern __shared__ uint32_t smem_raw[];
		constexpr int E = L + QB_EXT_LIMBS;
		const int nk = int(cx.lambda) + 1;
		const int g = threadIdx.x / nk, k = threadIdx.x % nk;
		const int t = blockIdx.x * G + g;
		const bool active = (g < G) && (t < n_tables);
		mpf<L>* buf0 = reinterpret_cast<mpf<L>*>(smem_raw);
		mpf<L>* buf1 = buf0 + G * nk;
		const int prec = cx.prec;
		const uint32_t n_max = cx.n_max;

		mpf<L> rho, q, qprev, s, a, ap, c4, rp, tmp;
		copy(rho, *reinterpret_cast<const mpf<L>*>(cx.rho));
		const mpf<L>* b0p = nullptr;
		const mpf<L>* b1p = nullptr;
		set_zero(s);
		set_zero(c4);
		set_zero(rp);
		set_zero(q);
		set_zero(qprev);
		if (active)
		{
			// q_k = Delta - k by k rounded decrements                      (real_function.hpp:239 `pow_ -= 1`)
			copy(q, delta[t]);
			copy(qprev, q);
			for (int j = 0; j < k; ++j)
			{
				copy(qprev, q);
				add_si(q, q, -1, prec);
			}
			TableRef tr = tref[t];
			b0p = b + size_t(tr.job0) * (n_max + 1);
			b1p = tr.job1 >= 0 ? b + size_t(tr.job1) * (n_max + 1) : nullptr;
			if (k == 0)
				pow_from_log<L>(c4, *reinterpret_cast<const mpf<E>*>(cx.ln4e), q, *reinterpret_cast<const mpf<E>*>(cx.ln2e),
				                prec);  // 4^Delta
			else
				pow_from_log<L>(rp, *reinterpret_cast<const mpf<E>*>(cx.lnrhoe), q, *reinterpret_cast<const mpf<E>*>(cx.ln2e),
				                prec);  // rho^(q_k)
		}
		const int steps = int(n_max) + nk;
		for (int tau = 0; tau < steps; ++tau)
		{
			mpf<L>* cur = (tau & 1) ? buf1 : buf0;
			mpf<L>* prv = (tau & 1) ? buf0 : buf1;
			int idx = tau - k;  // 0 .. n_max  <->  n = n_max - idx
			if (active && idx >= 0 && idx <= int(n_max))
			{
				uint32_t n = n_max - uint32_t(idx);
				if (k == 0)
				{
					copy(ap, b0p[n]);
					if (b1p)
					{
						copy(tmp, b1p[n]);
						add(ap, ap, tmp, prec);
						mul_2si(ap, ap, -1);
					}
					mul(a, ap, c4, prec);                                      // h *= 4^Delta (hor_recursion.cpp:43)
				}
				else
				{
					copy(ap, prv[g * nk + k - 1]);
					deriv_scale(a, ap, qprev, n, prec);
				}
				if (idx == 0)
					copy(s, a);
				else
					deriv_horner(s, rho, a, prec);
				copy(cur[g * nk + k], a);
			}
			__syncthreads();
		}
		if (active)
		{
			if (k == 0)
				pow_from_log<L>(rp, *reinterpret_cast<const mpf<E>*>(cx.lnrhoe), q, *reinterpret_cast<const mpf<E>*>(cx.ln2e),
				                prec);
			mpf<L> kf, f;
			copy(kf, reinterpret_cast<const mpf<L>*>(cx.kfact)[k]);
			deriv_finish(f, s, rp, kf, uint32_t(k), prec);
			copy(fr[size_t(t) * nk + k], f);
		}
	}

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

	template <int L>
	__global__ void k_vtod(DevCtx cx, int n_d, const mpf<L>* __restrict__ d, mpf<L>* __restrict__ v)
	{
		int i = blockIdx.x * blockDim.x + threadIdx.x;
		if (i >= n_d) return;
		constexpr int E = L + QB_EXT_LIMBS;
		mpf<L> md, p4;
		copy(md, d[i]);
		neg(md);
		pow_from_log<L>(p4, *reinterpret_cast<const mpf<E>*>(cx.ln4e), md, *reinterpret_cast<const mpf<E>*>(cx.ln2e), cx.prec);
		v_to_d_table(v + size_t(i) * cf_dim(int(cx.lambda)), 1, int(cx.lambda), d[i], p4, cx.prec);
	}

	// term = (table index, d index, sym); output packed per term at out[out_ofs ...]
	struct TermRef
	{
		int32_t table, dix, sym, out_ofs;
	};
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

	// Elementwise primitive (diagnostic entry qb_op_batch): lets the GPU test-suite pin the device build of every
	// arithmetic routine against MPFR.  op codes follow oracle/ref_capi.cpp.
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
		case 23: r.sk = uint32_t(cmp(x, y) + 1); break;
		case 24: r.sk = is_integer(x) ? 1u : 0u; break;
		default: set_nan(r); break;
		}
		copy(out[i], r);
	}

	// Integer multiply-add issue-rate probe: 8 independent mad.wide.u32 accumulation chains per thread, no memory
	// traffic.  One "limb-MAC" (32 x 32 -> 64 multiply plus 64-bit accumulate) is the unit of the roofline.
	static __global__ void k_imad_peak(uint64_t* out, int iters, uint32_t seed)
	{
		uint32_t a = seed + threadIdx.x, b = seed * 2654435761u + blockIdx.x;
		uint64_t c0 = 1, c1 = 2, c2 = 3, c3 = 4, c4 = 5, c5 = 6, c6 = 7, c7 = 8;
#pragma unroll 1
		for (int i = 0; i < iters; ++i)
		{
#pragma unroll
			for (int u = 0; u < 8; ++u)
			{
				asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c0) : "r"(a), "r"(b));
				asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c1) : "r"(a), "r"(b));
				asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c2) : "r"(a), "r"(b));
				asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c3) : "r"(a), "r"(b));
				asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c4) : "r"(a), "r"(b));
				asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c5) : "r"(a), "r"(b));
				asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c6) : "r"(a), "r"(b));
				asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c7) : "r"(a), "r"(b));
			}
		}
		uint64_t s = c0 ^ c1 ^ c2 ^ c3 ^ c4 ^ c5 ^ c6 ^ c7;
		if (s == 0x1234567ull) out[0] = s;  // keep the chains alive
	}
}  // namespace qb
