// qboot_b200 -- CUDA kernels of the block-table pipeline (sm_100a), templated on the limb count L.
//
//   k_rec_coeffs   one thread per (series job, recurrence coefficient)        -> coef[job][40]
//   k_polyvals     one thread per (series job, n): the recurrence polynomials at n  -> pv[job][n][8]
//   k_series       one thread per series job, n_max sequential steps          -> b[job][0..n_max]
//   k_pow          one thread per (table, power): 4^Delta and rho^(Delta-k)              -> pw[table][lambda+2]
//   k_scale        one thread per (table, n): the lambda+1 scaled coefficient arrays     -> a[table][k][n]
//   k_horner       one thread per (table, k): Horner sum, times rho^q / k!               -> fr[table][k]
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
#include "launch.cuh"
#include "hor_table.inc"

namespace qb
{
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
		const int K = sp == 0 ? 6 : 8, deg = sp == 0 ? 3 : 4, st = deg + 1;
		const mpf<L>* c = coef + size_t(job) * QB_NCOEF;
		mpf<L>* out = pv + size_t(gid) * 8;
		for (int i = 0; i < K; ++i)
		{
			mpf<L> s;
			copy(s, c[i * st + deg]);
			for (int d = deg - 1; d >= 0; --d)
			{
				fast::mul_small(s, s, int64_t(n), cx.prec);
				fast::add(s, s, c[i * st + d], cx.prec);
			}
			copy(out[i], s);
		}
	}

	// b[n] = -(sum_{i=1}^{min(n,K-1)} p_i(n) b[n-i]) / p_0(n), one thread per series job (src/hor_recursion.cpp:31-37).
	// The running sum stays in registers; p_i(n) and the previous terms are read where they lie in global memory.
	template <int L>
	__global__ void k_series(DevCtx cx, int n_jobs, const mpf<L>* __restrict__ delta, const uint32_t* __restrict__ spin,
	                         const mpf<L>* __restrict__ b0, const mpf<L>* __restrict__ pv, mpf<L>* __restrict__ b)
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
		const int K = spin[job] == 0 ? 6 : 8;
		const int prec = cx.prec;
		const mpf<L>* pj = pv + size_t(job) * cx.n_max * 8;
		for (uint32_t n = 1; n <= cx.n_max; ++n)
		{
			const mpf<L>* p = pj + size_t(n - 1) * 8;
			mpf<L> sum, bn;
			fast::mul(sum, p[1], bj[n - 1], prec);  // fma(0, p_1 b[n-1]) rounds like the product
			for (int i = 2; i < K; ++i)
				if (uint32_t(i) <= n) fast::fma(sum, p[i], bj[n - uint32_t(i)], sum, prec);
			neg(sum);
			fast::div(bn, sum, p[0], prec);
			copy(bj[n], bn);
		}
	}

	// powers of one table: pw[t][0] = 4^Delta (src/hor_recursion.cpp:43), pw[t][1 + k] = rho^(q_k) with q_k = Delta - k by
	// k rounded decrements (include/qboot/algebra/real_function.hpp:239,262); one thread per power
	template <int L>
	__global__ void k_pow(DevCtx cx, int n_tables, const mpf<L>* __restrict__ delta, mpf<L>* __restrict__ pw)
	{
		constexpr int E = L + QB_EXT_LIMBS;
		const int np = int(cx.lambda) + 2;
		int gid = blockIdx.x * blockDim.x + threadIdx.x;
		int t = gid / np, j = gid % np;
		if (t >= n_tables) return;
		mpf<L> q, r;
		copy(q, delta[t]);
		for (int i = 1; i < j; ++i) add_si(q, q, -1, cx.prec);
		if (j == 0)
			pow_from_log<L>(r, *reinterpret_cast<const mpf<E>*>(cx.ln4e), q, *reinterpret_cast<const mpf<E>*>(cx.ln2e), cx.prec);
		else
			pow_from_log<L>(r, *reinterpret_cast<const mpf<E>*>(cx.lnrhoe), q, *reinterpret_cast<const mpf<E>*>(cx.ln2e), cx.prec);
		copy(pw[size_t(t) * np + j], r);
	}

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
