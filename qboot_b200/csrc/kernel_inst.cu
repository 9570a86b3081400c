// qboot_b200 -- kernel translation unit: nvcc -DQB_L=<limbs> -DQB_KGROUP=<0..4> kernel_inst.cu
//   group 0: k_rec_coeffs, k_polyvals, k_ops    1: k_series    2: k_deriv    3: k_finish    4: k_vtod, k_fblock, k_pow
#if !defined(QB_L) || !defined(QB_KGROUP)
#error "compile with -DQB_L=<number of 32-bit limbs> -DQB_KGROUP=<0..4>"
#endif
#include "kernels.cuh"

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
	void launch_polyvals<QB_L>(const DevCtx& cx, int n_jobs, const uint32_t* spin, const mpf<QB_L>* coef, mpf<QB_L>* pv,
	                           cudaStream_t st)
	{
		int threads = 128;
		int64_t total = int64_t(n_jobs) * cx.n_max;
		k_polyvals<QB_L><<<unsigned((total + threads - 1) / threads), threads, 0, st>>>(cx, n_jobs, spin, coef, pv);
	}
	template <>
	void launch_ops<QB_L>(const DevCtx& cx, int op, int n, const mpf<QB_L>* a, const mpf<QB_L>* b, const mpf<QB_L>* c,
	                      const int64_t* ia, const int64_t* ib, mpf<QB_L>* out, cudaStream_t st)
	{
		k_ops<QB_L><<<(n + 63) / 64, 64, 0, st>>>(cx, op, n, a, b, c, ia, ib, out);
	}
#elif QB_KGROUP == 1
	template <>
	void launch_series<QB_L>(const DevCtx& cx, int n_jobs, const mpf<QB_L>* delta, const uint32_t* spin, const mpf<QB_L>* b0,
	                         const mpf<QB_L>* pv, mpf<QB_L>* b, cudaStream_t st)
	{
		int threads = 32;
		k_series<QB_L><<<(n_jobs + threads - 1) / threads, threads, 0, st>>>(cx, n_jobs, delta, spin, b0, pv, b);
	}
#elif QB_KGROUP == 2
	template <>
	cudaError_t launch_deriv<QB_L>(const DevCtx& cx, int n_tables, const TableRef* tref, const mpf<QB_L>* delta,
	                               const mpf<QB_L>* b, const mpf<QB_L>* pw, mpf<QB_L>* fr, cudaStream_t st)
	{
		const int nk = int(cx.lambda) + 1;
		using SR = SRec<QB_L>;
		int G = 160 / nk;
		if (G < 1) G = 1;
		auto bytes = [&](int g) { return size_t(4 * g * nk + g + 1) * SR::SW * sizeof(uint32_t); };
		while (bytes(G) > 110 * 1024 && G > 1) --G;   // two CTAs per SM
		size_t smem = bytes(G);
		if (smem > 48 * 1024)
		{
			cudaError_t e = cudaFuncSetAttribute(k_deriv<QB_L>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
			if (e != cudaSuccess) return e;
		}
		int blocks = (n_tables + G - 1) / G;
		k_deriv<QB_L><<<blocks, G * nk, smem, st>>>(cx, n_tables, G, tref, delta, b, pw, fr);
		return cudaSuccess;
	}
#elif QB_KGROUP == 3
	template <>
	void launch_finish<QB_L>(const DevCtx& cx, int n_tables, const mpf<QB_L>* delta, const uint32_t* spin, const mpf<QB_L>* S,
	                         const mpf<QB_L>* P, const mpf<QB_L>* fr, mpf<QB_L>* g, cudaStream_t st)
	{
		int threads = 64;
		k_finish<QB_L><<<(n_tables + threads - 1) / threads, threads, 0, st>>>(cx, n_tables, delta, spin, S, P, fr, g);
	}
#elif QB_KGROUP == 4
	template <>
	void launch_pow<QB_L>(const DevCtx& cx, int n_tables, const mpf<QB_L>* delta, mpf<QB_L>* pw, cudaStream_t st)
	{
		int threads = 128;
		int64_t total = int64_t(n_tables) * (int(cx.lambda) + 2);
		k_pow<QB_L><<<unsigned((total + threads - 1) / threads), threads, 0, st>>>(cx, n_tables, delta, pw);
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
#else
#error "QB_KGROUP out of range"
#endif
}  // namespace qb
