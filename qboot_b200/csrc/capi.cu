// qboot_b200 -- C ABI (include/qboot_b200.h) over the per-precision engines.
#include <cuda_runtime.h>

#include <cstring>
#include <new>

#include "../../include/qboot_b200.h"
#include "engine.cuh"

#ifndef QB_L_LIST
#error "QB_L_LIST must list the compiled limb counts, e.g. -DQB_L_LIST='X(19) X(25) X(32) X(48)'"
#endif

#define X(LL) extern "C" qb::EngineBase* qb_new_engine_L##LL(int device, int prec);
QB_L_LIST
#undef X

namespace qb
{
	// Limb-product issue-rate probe: 16 independent 32 x 32 + 64 -> 64 multiply-accumulate chains per thread
	// (IMAD.WIDE.U32), no memory traffic.  One "limb-MAC" is the unit of the roofline.  The multiplier changes every
	// iteration and differs per chain: with loop-invariant operands ptxas hoists the product and the loop degenerates
	// into additions, which measures the wrong pipe (it did, in the first version of this probe).
	static __global__ void k_imad_peak(uint64_t* out, int iters, uint32_t seed)
	{
		uint32_t a[16], y = seed * 2654435761u + blockIdx.x;
		uint64_t c[16];
#pragma unroll
		for (int u = 0; u < 16; ++u)
		{
			a[u] = (seed + threadIdx.x) * uint32_t(2 * u + 3) + 1u;
			c[u] = u + 1;
		}
#pragma unroll 1
		for (int i = 0; i < iters; ++i)
		{
#pragma unroll
			for (int r = 0; r < 4; ++r)
			{
#pragma unroll
				for (int u = 0; u < 16; ++u) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c[u]) : "r"(a[u]), "r"(y));
				y += 0x9e3779b9u;
			}
		}
		uint64_t s = 0;
#pragma unroll
		for (int u = 0; u < 16; ++u) s ^= c[u];
		if (s == 0x1234567ull) out[0] = s;  // keep the chains alive
	}
	void launch_imad_peak(uint64_t* out, int iters, uint32_t seed, int blocks, int threads, cudaStream_t st)
	{
		k_imad_peak<<<blocks, threads, 0, st>>>(out, iters, seed);
	}
}  // namespace qb

struct qb_handle
{
	qb::EngineBase* e;
};

// nothing may throw across the C ABI: host-side planning allocates (std::vector, std::thread)
#define QB_GUARDED(expr)               \
	try                                \
	{                                  \
		return (expr);                 \
	}                                  \
	catch (const std::bad_alloc&)      \
	{                                  \
		return QB_ERR_NO_MEMORY;       \
	}                                  \
	catch (...)                        \
	{                                  \
		return QB_ERR_BAD_ARG;         \
	}

static qb::EngineFactory factory_for(int L)
{
	switch (L)
	{
#define X(LL) \
	case LL: return &qb_new_engine_L##LL;
		QB_L_LIST
#undef X
	default: return nullptr;
	}
}

extern "C"
{
	const char* qb_version(void) { return "qboot_b200 0.1 (sm_100a)"; }
	const char* qb_strerror(int s)
	{
		switch (s)
		{
		case QB_OK: return "ok";
		case QB_ERR_NO_DEVICE: return "no usable CUDA device (this engine has no CPU fallback)";
		case QB_ERR_BAD_PREC: return "no kernels compiled for this precision";
		case QB_ERR_NO_CONTEXT: return "qb_context_init has not been called on this handle";
		case QB_ERR_BAD_ARG: return "bad argument";
		case QB_ERR_CUDA: return "CUDA error (see qb_last_cuda_error)";
		case QB_ERR_RANGE: return "rational dimension or spin outside the supported range";
		case QB_ERR_ROUNDING: return "a power could not be rounded with certainty at the second Ziv level (see qb_ziv_counters)";
		case QB_ERR_NO_MEMORY: return "device memory exhausted";
		default: return "unknown status";
		}
	}
	int qb_precision_supported(int prec) { return prec > 0 && factory_for((prec + 31) / 32) != nullptr; }

	int qb_create(int device, int prec, qb_handle** out)
	{
		if (!out) return QB_ERR_BAD_ARG;
		*out = nullptr;
		if (prec < 64) return QB_ERR_BAD_PREC;
		qb::EngineFactory f = factory_for((prec + 31) / 32);
		if (!f) return QB_ERR_BAD_PREC;
		int ndev = 0;
		if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev)
		{
			cudaGetLastError();
			return QB_ERR_NO_DEVICE;
		}
		qb::EngineBase* e = f(device, prec);
		if (!e) return QB_ERR_NO_DEVICE;
		qb_handle* h = new (std::nothrow) qb_handle{e};
		if (!h)
		{
			delete e;
			return QB_ERR_NO_MEMORY;
		}
		*out = h;
		return QB_OK;
	}
	void qb_destroy(qb_handle* h)
	{
		if (!h) return;
		delete h->e;
		delete h;
	}
	int qb_words(const qb_handle* h) { return h ? h->e->limbs + 2 : QB_ERR_BAD_ARG; }
	const char* qb_last_cuda_error(const qb_handle* h) { return h ? h->e->cuda_error.c_str() : ""; }
	int qb_context_init(qb_handle* h, uint32_t n_max, uint32_t lambda, int64_t en, uint32_t ed)
	{
		QB_GUARDED(h ? h->e->context_init(n_max, lambda, en, ed) : QB_ERR_BAD_ARG)
	}
	int qb_context_get(const qb_handle* h, uint32_t* rho, uint32_t* mat)
	{
		return h ? h->e->context_get(rho, mat) : QB_ERR_BAD_ARG;
	}
	int qb_table_dim(const qb_handle* h) { return h ? h->e->table_dim() : QB_ERR_BAD_ARG; }
	int qb_block_dim(const qb_handle* h, int sym)
	{
		return (h && (sym == 0 || sym == 1)) ? h->e->block_dim(sym) : QB_ERR_BAD_ARG;
	}
	int qb_gblock_plan(qb_handle* h, int n, const uint32_t* spin, const uint32_t* delta, const uint32_t* S,
	                   const uint32_t* P)
	{
		QB_GUARDED(h ? h->e->gblock_plan(n, spin, delta, S, P) : QB_ERR_BAD_ARG)
	}
	int qb_gblock_run(qb_handle* h) { QB_GUARDED(h ? h->e->gblock_run() : QB_ERR_BAD_ARG) }
	int qb_gblock_fetch(qb_handle* h, uint32_t* out) { return (h && out) ? h->e->gblock_fetch(out) : QB_ERR_BAD_ARG; }
	int qb_sync(qb_handle* h)
	{
		if (!h) return QB_ERR_BAD_ARG;
		cudaSetDevice(h->e->device);
		cudaError_t e = cudaStreamSynchronize(h->e->stream);
		if (e != cudaSuccess)
		{
			h->e->cuda_error = std::string("cudaStreamSynchronize: ") + cudaGetErrorString(e);
			return QB_ERR_CUDA;
		}
		return QB_OK;
	}
	int qb_set_timing(qb_handle* h, int on)
	{
		if (!h) return QB_ERR_BAD_ARG;
		h->e->timing = on != 0;
		return QB_OK;
	}
	int qb_kernel_times(qb_handle* h, float* ms4)
	{
		if (!h || !ms4) return QB_ERR_BAD_ARG;
		int rc = qb_sync(h);
		if (rc) return rc;
		h->e->collect_times_public();
		for (int i = 0; i < 4; ++i) ms4[i] = h->e->kernel_ms[i];
		return QB_OK;
	}
	int qb_imad_peak(qb_handle* h, int iters, double* macs_per_s)
	{
		return (h && macs_per_s && iters > 0) ? h->e->imad_peak(iters, macs_per_s) : QB_ERR_BAD_ARG;
	}
	const uint32_t* qb_gblock_device_tables(const qb_handle* h) { return h ? h->e->device_tables() : nullptr; }
	int64_t qb_launch_count(const qb_handle* h) { return h ? h->e->launches : 0; }
	uint64_t qb_stream(const qb_handle* h) { return h ? uint64_t(reinterpret_cast<uintptr_t>(h->e->stream)) : 0; }
	int qb_gblock_batch(qb_handle* h, int n, const uint32_t* spin, const uint32_t* delta, const uint32_t* S,
	                    const uint32_t* P, uint32_t* out)
	{
		if (!h) return QB_ERR_BAD_ARG;
		if (n > 0 && !out) return QB_ERR_BAD_ARG;
		int rc = qb_gblock_plan(h, n, spin, delta, S, P);
		if (rc) return rc;
		// every chunk is copied out as soon as it is finished, underneath the later chunks
		h->e->fetch_dst = out;
		rc = qb_gblock_run(h);
		h->e->fetch_dst = nullptr;
		if (rc)
		{
			// drain whatever was queued before the failure: copies into `out` must not outlive the call
			cudaSetDevice(h->e->device);
			cudaDeviceSynchronize();
			cudaGetLastError();
			return rc;
		}
		rc = qb_sync(h);
		return rc ? rc : h->e->check_rounding();
	}
	// ---- multi-GPU: tables of every rank gathered in one rank's HBM by peer copies -------------------------------------
	int qb_gblock_set_sink(qb_handle* h, void* dst)
	{
		if (!h) return QB_ERR_BAD_ARG;
		h->e->sink = static_cast<uint32_t*>(dst);
		return QB_OK;
	}
	void* qb_device_alloc(qb_handle* h, size_t bytes)
	{
		void* p = nullptr;
		if (!h || bytes == 0) return nullptr;
		if (cudaSetDevice(h->e->device) != cudaSuccess || cudaMalloc(&p, bytes) != cudaSuccess)
		{
			cudaGetLastError();
			return nullptr;
		}
		return p;
	}
	void qb_device_free(qb_handle* h, void* p)
	{
		if (!h || !p) return;
		cudaSetDevice(h->e->device);
		cudaFree(p);
	}
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "qb_ipc_export writes a 64-byte handle");
	int qb_ipc_export(qb_handle* h, const void* dev_ptr, unsigned char* handle64)
	{
		if (!h || !dev_ptr || !handle64) return QB_ERR_BAD_ARG;
		cudaIpcMemHandle_t hd;
		cudaSetDevice(h->e->device);
		cudaError_t e = cudaIpcGetMemHandle(&hd, const_cast<void*>(dev_ptr));
		if (e != cudaSuccess)
		{
			h->e->cuda_error = std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e);
			cudaGetLastError();
			return QB_ERR_CUDA;
		}
		std::memcpy(handle64, &hd, 64);
		return QB_OK;
	}
	int qb_ipc_open(qb_handle* h, const unsigned char* handle64, void** dev_ptr)
	{
		if (!h || !handle64 || !dev_ptr) return QB_ERR_BAD_ARG;
		cudaIpcMemHandle_t hd;
		std::memcpy(&hd, handle64, 64);
		cudaSetDevice(h->e->device);
		cudaError_t e = cudaIpcOpenMemHandle(dev_ptr, hd, cudaIpcMemLazyEnablePeerAccess);
		if (e != cudaSuccess)
		{
			h->e->cuda_error = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e);
			cudaGetLastError();
			return QB_ERR_CUDA;
		}
		return QB_OK;
	}
	int qb_ipc_close(qb_handle* h, void* dev_ptr)
	{
		if (!h || !dev_ptr) return QB_ERR_BAD_ARG;
		cudaSetDevice(h->e->device);
		cudaError_t e = cudaIpcCloseMemHandle(dev_ptr);
		if (e != cudaSuccess)
		{
			cudaGetLastError();
			return QB_ERR_CUDA;
		}
		return QB_OK;
	}
	int qb_set_ziv_slack(qb_handle* h, int bits) { return h ? h->e->set_ziv_slack(bits) : QB_ERR_BAD_ARG; }
	int qb_ziv_counters(qb_handle* h, uint64_t* out2) { QB_GUARDED((h && out2) ? h->e->ziv_counters(out2) : QB_ERR_BAD_ARG) }
	int qb_op_batch(qb_handle* h, int op, int n, const uint32_t* a, const uint32_t* b, const uint32_t* c,
	                const int64_t* ia, const int64_t* ib, uint32_t* out)
	{
		return h ? h->e->op_batch(op, n, a, b, c, ia, ib, out) : QB_ERR_BAD_ARG;
	}
	int qb_fblock_batch(qb_handle* h, int m, const int32_t* table_idx, int n_d, const uint32_t* d, const int32_t* d_idx,
	                    const int32_t* sym, uint32_t* out)
	{
		QB_GUARDED(h ? h->e->fblock_batch(m, table_idx, n_d, d, d_idx, sym, out) : QB_ERR_BAD_ARG)
	}
	int qb_vtod_batch(qb_handle* h, int n_d, const uint32_t* d, uint32_t* out) { QB_GUARDED(h ? h->e->vtod_batch(n_d, d, out) : QB_ERR_BAD_ARG) }
	int qb_fblock_tables(qb_handle* h, int m, const uint32_t* g, const uint32_t* d, const int32_t* sym, uint32_t* out)
	{
		QB_GUARDED(h ? h->e->fblock_tables(m, g, d, sym, out) : QB_ERR_BAD_ARG)
	}
	void* qb_pinned_alloc(size_t bytes)
	{
		void* p = nullptr;
		if (bytes == 0) return nullptr;
		if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess)
		{
			cudaGetLastError();
			return nullptr;
		}
		return p;
	}
	void qb_pinned_free(void* p)
	{
		if (p) cudaFreeHost(p);
	}
	// ---- PMP stage ----
	static_assert(sizeof(qb_pmp_term) == 20, "qb_pmp_term layout: five int32, as qb::PmpTerm (pmp.cuh)");
	static_assert(sizeof(qb_pmp_block) == 16, "qb_pmp_block layout");
	int qb_pmp_assemble(qb_handle* h, int n_eq, const int32_t* eq_dim, const int32_t* eq_sym, int n_blocks,
	                    const qb_pmp_block* blocks, int n_terms, const qb_pmp_term* terms, int n_tab, const int32_t* tab,
	                    int n_d, const uint32_t* d, int n_coef, const uint32_t* coef)
	{
		QB_GUARDED(h ? h->e->pmp_assemble(n_eq, eq_dim, eq_sym, n_blocks, reinterpret_cast<const int32_t*>(blocks), n_terms,
		                                  reinterpret_cast<const int32_t*>(terms), n_tab, tab, n_d, d, n_coef, coef)
		             : QB_ERR_BAD_ARG)
	}
	int qb_pmp_block_add(qb_handle* h, int n_samples, int sz, const uint32_t* mat, const uint32_t* target)
	{
		QB_GUARDED(h ? h->e->pmp_block_add(n_samples, sz, mat, target) : QB_ERR_BAD_ARG)
	}
	int qb_pmp_block_fetch(qb_handle* h, int block, uint32_t* out) { QB_GUARDED(h ? h->e->pmp_block_fetch(block, out) : QB_ERR_BAD_ARG) }
	int qb_pmp_pack(qb_handle* h, int n_sel, const int32_t* sel, int n_elim, const int32_t* lead, int n_free,
	                const int32_t* free_idx, const uint32_t* eqn, const uint32_t* eqn_target)
	{
		QB_GUARDED(h ? h->e->pmp_pack(n_sel, sel, n_elim, lead, n_free, free_idx, eqn, eqn_target) : QB_ERR_BAD_ARG)
	}
	int qb_pmp_packed_fetch(qb_handle* h, int j, uint32_t* B, uint32_t* c) { QB_GUARDED(h ? h->e->pmp_packed_fetch(j, B, c) : QB_ERR_BAD_ARG) }
	int qb_pmp_text(qb_handle* h, int ndigits, int64_t* sizes) { QB_GUARDED(h ? h->e->pmp_text(ndigits, sizes) : QB_ERR_BAD_ARG) }
	int qb_pmp_text_declined(qb_handle* h, int cap, int64_t* idx, uint32_t* recs)
	{
		QB_GUARDED((h && idx && recs && cap >= 0) ? h->e->pmp_text_declined(cap, idx, recs) : QB_ERR_BAD_ARG)
	}
	int qb_pmp_text_patch(qb_handle* h, int cnt, const int64_t* idx, const char* texts, int stride, const int32_t* len, int64_t* sizes)
	{
		QB_GUARDED(h ? h->e->pmp_text_patch(cnt, idx, texts, stride, len, sizes) : QB_ERR_BAD_ARG)
	}
	int qb_pmp_text_fetch(qb_handle* h, char* out) { QB_GUARDED(h ? h->e->pmp_text_fetch(out) : QB_ERR_BAD_ARG) }
	int qb_pmp_info(const qb_handle* h, int64_t* out4) { return (h && out4) ? h->e->pmp_info(out4) : QB_ERR_BAD_ARG; }
	static_assert(sizeof(qb_eval_item) == 12, "qb_eval_item layout: three int32, as qb::EvalItem (pmp_types.cuh)");
	int qb_poly_eval_batch(qb_handle* h, int n_polys, const int32_t* poly_ofs, const uint32_t* coef, int n_x, const uint32_t* x, int n_s,
	                       const uint32_t* s, int n_items, const qb_eval_item* items, uint32_t* out)
	{
		QB_GUARDED(h ? h->e->poly_eval_batch(n_polys, poly_ofs, coef, n_x, x, n_s, s, n_items, reinterpret_cast<const int32_t*>(items), out)
		             : QB_ERR_BAD_ARG)
	}
	int qb_dec_format(qb_handle* h, int ndigits, int n, const uint32_t* x, char* slots, int slot, int32_t* len)
	{
		QB_GUARDED(h ? h->e->dec_format(ndigits, n, x, slots, slot, len) : QB_ERR_BAD_ARG)
	}
}
