// qboot_b200 -- kernel translation unit: nvcc -DQB_L=<limbs> -DQB_KGROUP=<0..7> kernel_inst.cu
//   group 0: k_rec_coeffs, k_ops    1: k_series    2: k_scale, k_horner    3: k_rho_to_z, k_offdiag_val, k_offdiag_close    4: k_vtod, k_fblock, k_pow
//   group 6: k_horner (its own unit so that k_scale and k_horner can differ in QB_COLD_COPY)
//   group 7: k_series_coop (small batches: eight warps per 32 jobs)
//   group 5: k_pmp_accum, k_pmp_pack, k_dec_format, k_dec_gather (the last one is not templated: compiled once, -DQB_L_FIRST=<smallest L>)
#if !defined(QB_L) || !defined(QB_KGROUP)
#error "compile with -DQB_L=<number of 32-bit limbs> -DQB_KGROUP=<0..7>"
#endif
#include <cstdlib>
#if (QB_KGROUP == 1 || QB_KGROUP == 2 || QB_KGROUP == 7) && !defined(QB_COLD_COPY)
#define QB_COLD_COPY 1  // see fastops.cuh, "cold paths": measured per group -- on for k_series (-19 %) and k_scale (-2 %), off elsewhere
#endif
#if QB_KGROUP == 0
#include "kernels_g0.cuh"
#elif QB_KGROUP == 1
#include "kernels_g1.cuh"
#elif QB_KGROUP == 2 || QB_KGROUP == 6
#include "kernels_g2.cuh"
#elif QB_KGROUP == 3
#include "kernels_g3.cuh"
#elif QB_KGROUP == 4
#include "kernels_g4.cuh"
#elif QB_KGROUP == 7
#include "kernels_g7.cuh"
#else
#include "kernels_g5.cuh"
#endif

namespace qb
{
#if QB_KGROUP == 0
	template <>
	void launch_rec_coeffs<QB_L>(const DevCtx& cx, int n_jobs, const mpf<QB_L>* delta, const uint32_t* spin, const mpf<QB_L>* S,
	                             const mpf<QB_L>* P, mpf<QB_L>* coef, cudaStream_t st)
	{
		int threads = 128, total = n_jobs * QB_NCOEF;
		k_rec_coeffs<QB_L><<<(total + threads - 1) / threads, threads, 0, st>>>(cx, n_jobs, delta, spin, S, P, coef);
	}
	template <>
	void launch_ops<QB_L>(const DevCtx& cx, int op, int n, const mpf<QB_L>* a, const mpf<QB_L>* b, const mpf<QB_L>* c,
	                      const int64_t* ia, const int64_t* ib, mpf<QB_L>* out, cudaStream_t st)
	{
		k_ops<QB_L><<<(n + 63) / 64, 64, 0, st>>>(cx, op, n, a, b, c, ia, ib, out);
	}
#elif QB_KGROUP == 1
	template <>
	void launch_coef_fx<QB_L>(const DevCtx& cx, int n_jobs, const uint32_t* spin, const mpf<QB_L>* coef, uint32_t* coefx, cudaStream_t st)
	{
		const int threads = 128;
		const int64_t total = int64_t(n_jobs) * 8;
		k_coef_fx<QB_L><<<unsigned((total + threads - 1) / threads), threads, 0, st>>>(cx, n_jobs, spin, coef, coefx);
	}
	template <>
	void launch_series<QB_L>(const DevCtx& cx, int n_jobs, const mpf<QB_L>* delta, const uint32_t* spin, const mpf<QB_L>* b0,
	                         const mpf<QB_L>* coef, const uint32_t* coefx, mpf<QB_L>* b, cudaStream_t st)
	{
		static const bool attr_set = [] {
			cudaFuncSetAttribute(k_series<QB_L, QB_SERIES_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
			                     int(series_smem_bytes<QB_L, QB_SERIES_THREADS>()));
			cudaFuncSetAttribute(k_series<QB_L, QB_SERIES_THREADS_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
			                     int(series_smem_bytes<QB_L, QB_SERIES_THREADS_SMALL>()));
			return true;
		}();
		(void)attr_set;
		static const int sms = [] {
			int dev = 0, n = 148;
			cudaGetDevice(&dev);
			cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
			return n;
		}();
		// a batch that fits the cooperative kernel's single wave: eight warps per 32 jobs (kernels_g7.cuh); at most one small
		// CTA per SM: the small-CTA kernel (one warp per scheduler); otherwise the large lockstep CTAs
		// QB_SERIES_COOP: jobs per SM up to which the cooperative kernel runs (0 = never); read at every launch so that a test
		// can run one batch through both kernels
		const char* coop_env = std::getenv("QB_SERIES_COOP");
		const int coop_limit = coop_env ? std::atoi(coop_env) : 64;
		if (n_jobs <= sms * coop_limit)
			launch_series_coop<QB_L>(cx, n_jobs, delta, spin, b0, coef, coefx, b, st);
		else if (n_jobs <= sms * QB_SERIES_THREADS_SMALL)
			k_series<QB_L, QB_SERIES_THREADS_SMALL><<<(n_jobs + QB_SERIES_THREADS_SMALL - 1) / QB_SERIES_THREADS_SMALL, QB_SERIES_THREADS_SMALL,
			                                         series_smem_bytes<QB_L, QB_SERIES_THREADS_SMALL>(), st>>>(cx, n_jobs, delta, spin, b0, coef, coefx, b);
		else
			k_series<QB_L, QB_SERIES_THREADS><<<(n_jobs + QB_SERIES_THREADS - 1) / QB_SERIES_THREADS, QB_SERIES_THREADS,
			                                   series_smem_bytes<QB_L, QB_SERIES_THREADS>(), st>>>(cx, n_jobs, delta, spin, b0, coef, coefx, b);
	}
#elif QB_KGROUP == 2
	template <>
	void launch_scale<QB_L>(const DevCtx& cx, int n_tables, const TableRef* tref, const mpf<QB_L>* b, const mpf<QB_L>* pw,
	                        mpf<QB_L>* a, cudaStream_t st)
	{
		int threads = QB_DERIV_THREADS;
		int64_t total = int64_t(n_tables) * (cx.n_max + 1);
		k_scale<QB_L><<<unsigned((total + threads - 1) / threads), threads, 0, st>>>(cx, n_tables, tref, b, pw, a);
	}
#elif QB_KGROUP == 3
	template <>
	int launch_finish<QB_L>(const DevCtx& cx, int n_tables, const mpf<QB_L>* delta, const uint32_t* spin, const mpf<QB_L>* S,
	                        const mpf<QB_L>* P, const mpf<QB_L>* fr, mpf<QB_L>* g, mpf<QB_L>* qc, cudaStream_t st)
	{
		const int threads = 128, lambda = int(cx.lambda);
		auto blocks = [&](int64_t total) { return unsigned((total + threads - 1) / threads); };
		int count = 1;
		k_rho_to_z<QB_L><<<blocks(int64_t(n_tables) * ((lambda + 2) / 2)), threads, 0, st>>>(cx, n_tables, delta, spin, fr, g, qc);
		for (int n = 1; n <= lambda / 2; ++n)
		{
			k_offdiag_val<QB_L><<<unsigned((int64_t(n_tables) * (lambda - 2 * n + 1) + QB_OFFDIAG_THREADS - 1) / QB_OFFDIAG_THREADS), QB_OFFDIAG_THREADS, 0, st>>>(
			    cx, n_tables, n, S, P, qc, g);
			k_offdiag_close<QB_L><<<blocks(n_tables), threads, 0, st>>>(cx, n_tables, n, g);
			count += 2;
		}
		return count;
	}
#elif QB_KGROUP == 4
	template <>
	void launch_pow<QB_L>(const DevCtx& cx, int n_tables, const mpf<QB_L>* delta, mpf<QB_L>* pw, cudaStream_t st)
	{
		int threads = 32;
		k_pow<QB_L><<<(n_tables + threads - 1) / threads, threads, 0, st>>>(cx, n_tables, delta, pw);
	}
	template <>
	void launch_vtod<QB_L>(const DevCtx& cx, int n_d, const mpf<QB_L>* d, mpf<QB_L>* v, cudaStream_t st)
	{
		int threads = 32;
		k_vtod<QB_L><<<(n_d + threads - 1) / threads, threads, 0, st>>>(cx, n_d, d, v);
	}
	template <>
	void launch_fblock<QB_L>(const DevCtx& cx, int n_terms, int dim_max, const TermRef* terms, const mpf<QB_L>* g,
	                         const mpf<QB_L>* v, mpf<QB_L>* out, cudaStream_t st)
	{
		int threads = 128;
		int64_t total = int64_t(n_terms) * dim_max;
		k_fblock<QB_L><<<unsigned((total + threads - 1) / threads), threads, 0, st>>>(cx, n_terms, dim_max, terms, g, v, out);
	}
#elif QB_KGROUP == 5
	template <>
	void launch_pmp_accum<QB_L>(const DevCtx& cx, int n_blocks, const PmpBlock* blocks, const PmpTerm* terms, const int32_t* tab,
	                            int n_eq, const int32_t* eq_ofs, const int32_t* eq_sym, const mpf<QB_L>* g, const mpf<QB_L>* v,
	                            const mpf<QB_L>* coef, mpf<QB_L>* Marena, int64_t total, cudaStream_t st)
	{
		const int threads = 128;
		k_pmp_accum<QB_L><<<unsigned((total + threads - 1) / threads), threads, 0, st>>>(cx, n_blocks, blocks, terms, tab, n_eq, eq_ofs,
		                                                                                 eq_sym, g, v, coef, Marena, total);
	}
	template <>
	void launch_pmp_pack<QB_L>(const DevCtx& cx, int n_blocks, const PmpBlock* blocks, int N, int n_free, const int32_t* free_idx,
	                           int n_elim, const int32_t* lead, const mpf<QB_L>* eqn, const mpf<QB_L>* eqn_target,
	                           const mpf<QB_L>* Marena, const mpf<QB_L>* targets, mpf<QB_L>* Barena, mpf<QB_L>* Carena, int64_t total,
	                           cudaStream_t st)
	{
		const int threads = 128;
		k_pmp_pack<QB_L><<<unsigned((total + threads - 1) / threads), threads, 0, st>>>(cx, n_blocks, blocks, N, n_free, free_idx, n_elim,
		                                                                                lead, eqn, eqn_target, Marena, targets, Barena,
		                                                                                Carena, total);
	}
	template <>
	void launch_poly_eval<QB_L>(const DevCtx& cx, int n_items, const EvalItem* items, const int32_t* poly_ofs, const mpf<QB_L>* coef,
	                            const mpf<QB_L>* x, const mpf<QB_L>* s, mpf<QB_L>* out, cudaStream_t st)
	{
		const int threads = 128;
		k_poly_eval<QB_L><<<(n_items + threads - 1) / threads, threads, 0, st>>>(cx, n_items, items, poly_ofs, coef, x, s, out);
	}
	template <>
	void launch_dec_format<QB_L>(const DecTables& tb, int64_t n, const mpf<QB_L>* x, char* slots, int slot, int32_t* len, cudaStream_t st)
	{
		const int threads = 64;
		k_dec_format<QB_L><<<unsigned((n + threads - 1) / threads), threads, 0, st>>>(tb, n, x, slots, slot, len);
	}
#if QB_L == QB_L_FIRST
	void launch_dec_gather(int64_t n, const char* slots, int slot, const int32_t* len, const int64_t* ofs, char* text, cudaStream_t st)
	{
		const int threads = 256;
		k_dec_gather<<<unsigned((n * 32 + threads - 1) / threads), threads, 0, st>>>(n, slots, slot, len, ofs, text);
	}
#endif
#elif QB_KGROUP == 7
	template <>
	void launch_series_coop<QB_L>(const DevCtx& cx, int n_jobs, const mpf<QB_L>* delta, const uint32_t* spin, const mpf<QB_L>* b0,
	                              const mpf<QB_L>* coef, const uint32_t* coefx, mpf<QB_L>* b, cudaStream_t st)
	{
		static const bool attr_set = [] {
			cudaFuncSetAttribute(k_series_coop<QB_L>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(CoopSmem<QB_L>::bytes));
			return true;
		}();
		(void)attr_set;
		k_series_coop<QB_L><<<(n_jobs + QB_COOP_JOBS - 1) / QB_COOP_JOBS, QB_COOP_T, CoopSmem<QB_L>::bytes, st>>>(cx, n_jobs, delta, spin, b0,
		                                                                                                          coef, coefx, b);
	}
#elif QB_KGROUP == 6
	template <>
	void launch_horner<QB_L>(const DevCtx& cx, int n_tables, const mpf<QB_L>* a, const mpf<QB_L>* pw, mpf<QB_L>* fr,
	                         cudaStream_t st)
	{
		int threads = QB_DERIV_THREADS;
		int64_t total = int64_t(n_tables) * (int(cx.lambda) + 1);
		k_horner<QB_L><<<unsigned((total + threads - 1) / threads), threads, 0, st>>>(cx, n_tables, a, pw, fr);
	}
#else
#error "QB_KGROUP out of range"
#endif
}  // namespace qb
