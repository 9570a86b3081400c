// qboot_b200 -- launch interface between the per-precision engine (host code, engine.cuh) and the kernel translation
// units (kernel_inst.cu, one per limb count and kernel group so that the heavy device code compiles in parallel).
#pragma once

#include <cuda_runtime.h>

#include "hor.cuh"
#include "pmp_types.cuh"

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
		const uint32_t* rhoinv;   // rho^-k, k = 0..lambda, mpf<L + QB_EXT_LIMBS>
		// second Ziv level of the powers (mpf_math.cuh): the logarithms at L + QB_EXT2_LIMBS limbs, the width (in bits of the
		// extended result's last place) inside which a rounding counts as in doubt, and two device counters
		// ([0] second-level evaluations, [1] values still in doubt)
		const uint32_t* ln2w;
		const uint32_t* ln4w;
		const uint32_t* lnrhow;
		int ziv_slack;
		uint32_t* ziv;
	};

	// a series job: one run of the recurrence at (Delta', spin, S, P)
	// a table: references one job, or two for a divergent operator (averaged, src/hor_recursion.cpp:21-28)
	struct TableRef
	{
		int32_t job0, job1;  // job1 < 0: single
	};
	// F/H term = (table index, d index, sym); output packed per term at out[out_ofs ...]
	struct TermRef
	{
		int32_t table, dix, sym, out_ofs;
	};

	constexpr int QB_NCOEF = 40;  // 8 polynomials x 5 coefficients (spin 0 uses 6 x 4 of them)
	// The recurrence coefficients of one launch (n_jobs series jobs) are stored WORD-PLANAR across jobs: word w of
	// coefficient c of job j at base[(c * W + w) * n_jobs + j].  Writer (k_rec_coeffs) and reader (k_series) both run with
	// the job as the lane index, so every access of a warp is one contiguous 128-byte line (the record-per-job layout cost
	// 32 sectors per 8-byte load and is read 400 times per job).
	template <int L>
	QB_HD QB_INLINE void coef_store(uint32_t* base, int c, int job, int n_jobs, const mpf<L>& v)
	{
		uint32_t* p = base + size_t(c) * (L + 2) * size_t(n_jobs) + job;
		p[0] = v.sk;
		p[size_t(n_jobs)] = uint32_t(v.exp);
		for (int j = 0; j < L; ++j) p[size_t(2 + j) * size_t(n_jobs)] = v.m[j];
	}
	template <int L>
	QB_HD QB_INLINE void coef_load(mpf<L>& v, const uint32_t* base, int c, int job, int n_jobs)
	{
		const uint32_t* p = base + size_t(c) * (L + 2) * size_t(n_jobs) + job;
		v.sk = p[0];
		v.exp = int32_t(p[size_t(n_jobs)]);
		for (int j = 0; j < L; ++j) v.m[j] = p[size_t(2 + j) * size_t(n_jobs)];
	}

	template <int L>
	void launch_rec_coeffs(const DevCtx& cx, int n_jobs, const mpf<L>* delta, const uint32_t* spin, const mpf<L>* S,
	                       const mpf<L>* P, mpf<L>* coef, cudaStream_t st);
	// Fixed-point copy of the coefficients (polyfx.cuh), word-planar like `coef`: QB_FX_WORDS words per coefficient --
	// [0] sign | 1 if the whole polynomial is exact in its frame (meaningful in the leading coefficient), [1] frame exponent,
	// [2 .. 2 + L + 2) the magnitude in the frame.
	template <int L>
	QB_HD constexpr int fx_words()
	{
		return L + 4;
	}
	// Fixed-point copies of the recurrence coefficients (polyfx.cuh): blocks of 32 jobs, [block][coefficient][word][lane].
	// The words of one coefficient are 128 bytes apart and the coefficients of a polynomial follow each other, so a thread
	// walks a polynomial with compile-time offsets from one pointer, and a warp's access is one line.
	template <int L>
	QB_HD inline size_t coefx_offset(int job, int c)
	{
		return (size_t(job >> 5) * QB_NCOEF + size_t(c)) * size_t(fx_words<L>()) * 32 + size_t(job & 31);
	}
	template <int L>
	void launch_coef_fx(const DevCtx& cx, int n_jobs, const uint32_t* spin, const mpf<L>* coef, uint32_t* coefx, cudaStream_t st);
	template <int L>
	void launch_series(const DevCtx& cx, int n_jobs, const mpf<L>* delta, const uint32_t* spin, const mpf<L>* b0,
	                   const mpf<L>* coef, const uint32_t* coefx, mpf<L>* b, cudaStream_t st);
	template <int L>
	void launch_series_coop(const DevCtx& cx, int n_jobs, const mpf<L>* delta, const uint32_t* spin, const mpf<L>* b0,
	                        const mpf<L>* coef, const uint32_t* coefx, mpf<L>* b, cudaStream_t st);
	template <int L>
	void launch_scale(const DevCtx& cx, int n_tables, const TableRef* tref, const mpf<L>* b, const mpf<L>* pw, mpf<L>* a,
	                  cudaStream_t st);
	template <int L>
	void launch_horner(const DevCtx& cx, int n_tables, const mpf<L>* a, const mpf<L>* pw, mpf<L>* fr, cudaStream_t st);
	template <int L>
	void launch_pow(const DevCtx& cx, int n_tables, const mpf<L>* delta, mpf<L>* pw, cudaStream_t st);
	template <int L>
	int launch_finish(const DevCtx& cx, int n_tables, const mpf<L>* delta, const uint32_t* spin, const mpf<L>* S,
	                  const mpf<L>* P, const mpf<L>* fr, mpf<L>* g, mpf<L>* qc, cudaStream_t st);  // returns #launches
	template <int L>
	void launch_vtod(const DevCtx& cx, int n_d, const mpf<L>* d, mpf<L>* v, cudaStream_t st);
	template <int L>
	void launch_fblock(const DevCtx& cx, int n_terms, int dim_max, const TermRef* terms, const mpf<L>* g, const mpf<L>* v,
	                   mpf<L>* out, cudaStream_t st);
	template <int L>
	void launch_ops(const DevCtx& cx, int op, int n, const mpf<L>* a, const mpf<L>* b, const mpf<L>* c, const int64_t* ia,
	                const int64_t* ib, mpf<L>* out, cudaStream_t st);
	// PMP stage (kernels_g5.cuh)
	template <int L>
	void launch_pmp_accum(const DevCtx& cx, int n_blocks, const PmpBlock* blocks, const PmpTerm* terms, const int32_t* tab, int n_eq,
	                      const int32_t* eq_ofs, const int32_t* eq_sym, const mpf<L>* g, const mpf<L>* v, const mpf<L>* coef,
	                      mpf<L>* Marena, int64_t total, cudaStream_t st);
	template <int L>
	void launch_pmp_pack(const DevCtx& cx, int n_blocks, const PmpBlock* blocks, int N, int n_free, const int32_t* free_idx,
	                     int n_elim, const int32_t* lead, const mpf<L>* eqn, const mpf<L>* eqn_target, const mpf<L>* Marena,
	                     const mpf<L>* targets, mpf<L>* Barena, mpf<L>* Carena, int64_t total, cudaStream_t st);
	template <int L>
	void launch_poly_eval(const DevCtx& cx, int n_items, const EvalItem* items, const int32_t* poly_ofs, const mpf<L>* coef, const mpf<L>* x,
	                      const mpf<L>* s, mpf<L>* out, cudaStream_t st);
	template <int L>
	void launch_dec_format(const DecTables& tb, int64_t n, const mpf<L>* x, char* slots, int slot, int32_t* len, cudaStream_t st);
	void launch_dec_gather(int64_t n, const char* slots, int slot, const int32_t* len, const int64_t* ofs, char* text, cudaStream_t st);
	// integer multiply-add probe (capi.cu)
	void launch_imad_peak(uint64_t* out, int iters, uint32_t seed, int blocks, int threads, cudaStream_t st);
}  // namespace qb
