// qboot_b200 -- per-precision engine behind the C ABI (include/qboot_b200.h): device buffers, host-side planning and
// kernel launches.  One instantiation per limb count L lives in its own translation unit (engine_inst.cu, -DQB_L=..).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <map>
#include <new>
#include <string>
#include <vector>

namespace qb
{
	// precision-erased interface used by capi.cu
	struct EngineBase
	{
		int device = 0, prec = 0, limbs = 0;
		cudaStream_t stream = nullptr;
		std::string cuda_error;
		int64_t launches = 0;
		virtual ~EngineBase() {}
		virtual int context_init(uint32_t n_max, uint32_t lambda, int64_t eps_num, uint32_t eps_den) = 0;
		virtual int context_get(uint32_t* rho, uint32_t* mat) const = 0;
		virtual int table_dim() const = 0;
		virtual int block_dim(int sym) const = 0;
		virtual int gblock_plan(int n, const uint32_t* spin, const uint32_t* delta, const uint32_t* S, const uint32_t* P) = 0;
		virtual int gblock_run() = 0;
		virtual int gblock_fetch(uint32_t* out) = 0;
		virtual const uint32_t* device_tables() const = 0;
		virtual int fblock_batch(int m, const int32_t* table_idx, int n_d, const uint32_t* d, const int32_t* d_idx,
		                         const int32_t* sym, uint32_t* out) = 0;
		virtual int imad_peak(int iters, double* macs_per_s) = 0;
		virtual int collect_times_public() = 0;
		// per-kernel device times (ms) of the last gblock_run when timing is enabled: rec_coeffs, series, deriv, finish
		bool timing = false;
		float kernel_ms[4] = {0, 0, 0, 0};
		// when set, gblock_run copies every chunk's tables to this host buffer as soon as the chunk is finished (on a copy
		// stream, overlapping the later chunks); the handle's stream then also waits for the copies
		uint32_t* fetch_dst = nullptr;
		// the same for every later run, for callers that keep the tables on a device: host, device or PEER-device memory
		// (a pointer opened from another process's allocation, qb_ipc_open) -- the gather of a multi-GPU job
		uint32_t* sink = nullptr;
		virtual int op_batch(int op, int n, const uint32_t* a, const uint32_t* b, const uint32_t* c, const int64_t* ia,
		                     const int64_t* ib, uint32_t* out) = 0;
		virtual int vtod_batch(int n_d, const uint32_t* d, uint32_t* out) = 0;
		virtual int fblock_tables(int m, const uint32_t* g, const uint32_t* d, const int32_t* sym, uint32_t* out) = 0;
		// PMP stage (see include/qboot_b200.h, qb_pmp_*)
		virtual int pmp_assemble(int n_eq, const int32_t* eq_dim, const int32_t* eq_sym, int n_blocks, const int32_t* blocks4,
		                         int n_terms, const int32_t* terms5, int n_tab, const int32_t* tab, int n_d, const uint32_t* d,
		                         int n_coef, const uint32_t* coef) = 0;
		virtual int pmp_block_add(int n_samples, int sz, const uint32_t* mat, const uint32_t* target) = 0;
		virtual int pmp_block_fetch(int block, uint32_t* out) = 0;
		virtual int pmp_pack(int n_sel, const int32_t* sel, int n_elim, const int32_t* lead, int n_free, const int32_t* free_idx,
		                     const uint32_t* eqn, const uint32_t* eqn_target) = 0;
		virtual int pmp_packed_fetch(int j, uint32_t* B, uint32_t* c) = 0;
		virtual int pmp_text(int ndigits, int64_t* sizes) = 0;
		virtual int pmp_text_declined(int cap, int64_t* idx, uint32_t* recs) = 0;
		virtual int pmp_text_patch(int cnt, const int64_t* idx, const char* texts, int stride, const int32_t* len, int64_t* sizes) = 0;
		virtual int pmp_text_fetch(char* out) = 0;
		virtual int pmp_info(int64_t* out4) const = 0;
		virtual int dec_format(int ndigits, int n, const uint32_t* x, char* slots, int slot, int32_t* len) = 0;
		virtual int poly_eval_batch(int n_polys, const int32_t* poly_ofs, const uint32_t* coef, int n_x, const uint32_t* x, int n_s,
		                            const uint32_t* s, int n_items, const int32_t* items3, uint32_t* out) = 0;
		// Ziv loop of the powers: width of the in-doubt window (0 = default) and the counters [second-level evaluations,
		// values still in doubt] since the handle was created
		virtual int set_ziv_slack(int bits) = 0;
		virtual int ziv_counters(uint64_t* out2) = 0;
		virtual int check_rounding() = 0;
	};
	typedef EngineBase* (*EngineFactory)(int device, int prec);
}  // namespace qb

#ifdef QB_L
// ------------------------------------------------------------------------------------------------------------------
#include <chrono>
#include <cstdio>
#include <thread>
#include "hostmath.hpp"
#include "block.cuh"
#include "launch.cuh"

namespace qb
{
	// grow-only device buffer
	struct DevBuf
	{
		void* p = nullptr;
		size_t cap = 0;
		int ensure(size_t bytes)
		{
			if (bytes <= cap) return 0;
			if (p) cudaFree(p);
			p = nullptr;
			cap = 0;
			size_t want = bytes + bytes / 8 + 256;
			if (cudaMalloc(&p, want) != cudaSuccess)
			{
				cudaGetLastError();
				return -1;
			}
			cap = want;
			return 0;
		}
		// like ensure, but the first `keep` bytes survive a reallocation
		int grow(size_t bytes, size_t keep)
		{
			if (bytes <= cap) return 0;
			void* q = nullptr;
			size_t want = bytes + bytes / 4 + 256;
			if (cudaMalloc(&q, want) != cudaSuccess)
			{
				cudaGetLastError();
				return -1;
			}
			if (p && keep && cudaMemcpy(q, p, keep, cudaMemcpyDeviceToDevice) != cudaSuccess)
			{
				cudaGetLastError();
				cudaFree(q);
				return -1;
			}
			if (p) cudaFree(p);
			p = q;
			cap = want;
			return 0;
		}
		void release()
		{
			if (p) cudaFree(p);
			p = nullptr;
			cap = 0;
		}
		template <class T>
		T* as() const
		{
			return static_cast<T*>(p);
		}
	};

	template <int L>
	struct Engine : EngineBase
	{
		using T = mpf<L>;
		static constexpr int E = L + QB_EXT_LIMBS;
		bool has_ctx = false;
		uint32_t n_max = 0, lambda = 0;
		Rational eps{0, 1};
		ContextConsts<L> consts;
		std::map<uint32_t, T> b0_cache;
		DevBuf d_ziv;  // two counters of the powers' Ziv loop
		int ziv_slack = QB_ZIV_SLACK;
		uint64_t ziv_reported = 0;
		DevBuf d_b0tab;  // b[0] by spin (src/hor_recursion.cpp:17-19)
		DevBuf d_coefx;  // fixed-point copy of the recurrence coefficients (polyfx.cuh)
		DevBuf d_consts, d_in, d_jobs, d_coef, d_b, d_fr, d_g, d_tref, d_pw, d_qc, d_a, d_fb_in, d_fb_v, d_fb_out, d_fb_terms, d_fb_g;
		DevCtx cx{};
		int n_tables = 0, n_jobs = 0;
		bool tables_valid = false;  // a gblock_run has completed for the current plan
		// layout of d_in: [delta | S | P] (tables) then spin; d_jobs: [delta | S | P | b0] (jobs) then spin
		T *t_delta = nullptr, *t_S = nullptr, *t_P = nullptr;
		uint32_t* t_spin = nullptr;
		T *j_delta = nullptr, *j_S = nullptr, *j_P = nullptr;  // per job; alias the table arrays when no operator is divergent
		uint32_t* j_spin = nullptr;
		std::vector<cudaEvent_t> ev;    // timing: one in front of every stage (tick)
		std::vector<int> ev_stage;      // stage that follows ev[i], -1 none
		size_t ev_used = 0;
		std::vector<int> group_c;       // chunk boundaries of the front groups
		cudaStream_t lo_stream = nullptr, mid_stream = nullptr, fin_stream = nullptr;  // wide kernels of all chunks, by stage
		std::vector<cudaEvent_t> dep_ev;         // one per front group + two per chunk
		cudaEvent_t fork_ev = nullptr, copy_done_ev = nullptr, fin_done_ev = nullptr, lo_done_ev = nullptr;
		cudaStream_t copy_stream = nullptr;      // device -> host copies of finished chunks (fetch_dst)
		// side stream of a single-chunk run (one PMP): work that does not depend on the series recurrence -- the powers of
		// rho, the v^d tables of the assembly -- runs beside k_series, which leaves most SMs idle at that size
		cudaStream_t aux_stream = nullptr;
		cudaEvent_t aux_fork_ev = nullptr, aux_done_ev = nullptr;
		std::vector<int> job_of_table;  // first job of every table (+ sentinel)
		std::vector<int> chunk_t;       // table boundaries of the chunks
		int n_chunks = 1, sm_count = 0;
		size_t a_slot_records = 0;      // records of the scaled-coefficient scratch (largest chunk)

		Engine(int dev, int prec_bits)
		{
			device = dev;
			prec = prec_bits;
			limbs = L;
		}
		~Engine() override
		{
			cudaSetDevice(device);
			for (DevBuf* b : {&d_ziv, &d_b0tab, &d_coefx, &d_consts, &d_in, &d_jobs, &d_coef, &d_b, &d_fr, &d_g, &d_tref, &d_pw, &d_qc, &d_a, &d_fb_in, &d_fb_v, &d_fb_out,
			                  &d_fb_terms, &d_fb_g})
				b->release();
			pmp_release();
			for (cudaEvent_t e_ : ev) cudaEventDestroy(e_);
			for (cudaEvent_t e_ : dep_ev) cudaEventDestroy(e_);
			for (cudaEvent_t e_ : {fork_ev, copy_done_ev, fin_done_ev, lo_done_ev, aux_fork_ev, aux_done_ev})
				if (e_) cudaEventDestroy(e_);
			for (cudaStream_t s_ : {lo_stream, mid_stream, fin_stream, copy_stream, aux_stream})
				if (s_) cudaStreamDestroy(s_);
			if (stream) cudaStreamDestroy(stream);
		}
		// QB_DEBUG_SYNC=1: synchronise after every kernel so that a fault is attributed to the kernel that raised it
		int debug_sync(const char* kernel)
		{
			static const bool on = std::getenv("QB_DEBUG_SYNC") != nullptr;
			if (!on) return 0;
			cudaError_t e = cudaStreamSynchronize(stream);
			if (e == cudaSuccess) e = cudaGetLastError();
			if (e != cudaSuccess) return fail(e, kernel);
			return 0;
		}
		int fail(cudaError_t e, const char* where)
		{
			cuda_error = std::string(where) + ": " + cudaGetErrorString(e);
			cudaGetLastError();
			return -5;
		}
#define QB_CU(call)                                    \
	do                                                 \
	{                                                  \
		cudaError_t e_ = (call);                       \
		if (e_ != cudaSuccess) return fail(e_, #call); \
	} while (0)

		int ensure_aux()
		{
			if (aux_stream) return 0;
			QB_CU(cudaStreamCreateWithFlags(&aux_stream, cudaStreamNonBlocking));
			QB_CU(cudaEventCreateWithFlags(&aux_fork_ev, cudaEventDisableTiming));
			QB_CU(cudaEventCreateWithFlags(&aux_done_ev, cudaEventDisableTiming));
			return 0;
		}
		int context_init(uint32_t nmax, uint32_t lam, int64_t en, uint32_t ed) override
		{
			if (ed == 0 || lam > 200 || nmax == 0) return -4;
			{
				// the kernels hand small integers derived from epsilon to div_q / add_q / set_q (32-bit divisors): the transverse
				// recursion divides by 4 n eps + (4 n - 2) n (block.cuh), the Casimir uses 2 den; refuse what does not fit
				const int64_t an = en < 0 ? -en : en, nr = std::max<int64_t>(1, int64_t(lam) / 2);
				if (an >= (int64_t(1) << 40) || 2 * int64_t(ed) >= (int64_t(1) << 32)) return -6;
				const int64_t cn = 4 * nr * an + (4 * nr - 2) * nr * int64_t(ed);
				if (cn >= (int64_t(1) << 32)) return -6;
			}
			QB_CU(cudaSetDevice(device));
			// a new context invalidates the plan, the tables and the PMP blocks of the previous one
			n_tables = n_jobs = 0;
			tables_valid = false;
			pm_blocks.clear();
			pm_sel.clear();
			pm_N = 0;
			has_ctx = false;
			n_max = nmax;
			lambda = lam;
			// b[0] depends on (spin, epsilon) only: a scan that builds one Context per PMP keeps the values
			if (eps.num != en || eps.den != ed) b0_cache.clear();
			eps = Rational{en, ed};
			make_context_consts<L>(consts, lam, prec);
			// upload: rho | mat | kfact | ln2e | ln4e | lnrhoe
			size_t n1 = size_t(lam) + 1;
			constexpr int E2 = ContextConsts<L>::E2;
			const size_t ofs_w = (sizeof(T) * (1 + n1 * n1 + n1) + (3 + n1) * sizeof(mpf<E>) + 7) & ~size_t(7);
			size_t bytes = ofs_w + 3 * sizeof(mpf<E2>) + 8;
			if (d_consts.ensure(bytes)) return -7;
			std::vector<uint8_t> host(bytes);
			uint8_t* p = host.data();
			auto put = [&](const void* src, size_t n) {
				std::memcpy(p, src, n);
				p += n;
			};
			put(&consts.rho, sizeof(T));
			put(consts.mat.data(), sizeof(T) * n1 * n1);
			put(consts.kfact.data(), sizeof(T) * n1);
			put(&consts.ln2e, sizeof(mpf<E>));
			put(&consts.ln4e, sizeof(mpf<E>));
			put(&consts.lnrhoe, sizeof(mpf<E>));
			put(consts.rhoinv.data(), sizeof(mpf<E>) * n1);
			p = host.data() + ofs_w;
			put(&consts.ln2w, sizeof(mpf<E2>));
			put(&consts.ln4w, sizeof(mpf<E2>));
			put(&consts.lnrhow, sizeof(mpf<E2>));
			QB_CU(cudaMemcpyAsync(d_consts.p, host.data(), bytes, cudaMemcpyHostToDevice, stream));
			QB_CU(cudaStreamSynchronize(stream));
			const uint8_t* d = d_consts.as<uint8_t>();
			cx.prec = prec;
			cx.n_max = nmax;
			cx.lambda = lam;
			cx.eps = eps;
			cx.rho = reinterpret_cast<const uint32_t*>(d);
			cx.mat = reinterpret_cast<const uint32_t*>(d + sizeof(T));
			cx.kfact = reinterpret_cast<const uint32_t*>(d + sizeof(T) * (1 + n1 * n1));
			const uint8_t* de = d + sizeof(T) * (1 + n1 * n1 + n1);
			cx.ln2e = reinterpret_cast<const uint32_t*>(de);
			cx.ln4e = reinterpret_cast<const uint32_t*>(de + sizeof(mpf<E>));
			cx.lnrhoe = reinterpret_cast<const uint32_t*>(de + 2 * sizeof(mpf<E>));
			cx.rhoinv = reinterpret_cast<const uint32_t*>(de + 3 * sizeof(mpf<E>));
			cx.ln2w = reinterpret_cast<const uint32_t*>(d + ofs_w);
			cx.ln4w = reinterpret_cast<const uint32_t*>(d + ofs_w + sizeof(mpf<E2>));
			cx.lnrhow = reinterpret_cast<const uint32_t*>(d + ofs_w + 2 * sizeof(mpf<E2>));
			cx.ziv_slack = ziv_slack;
			if (!d_ziv.p)
			{
				if (d_ziv.ensure(8)) return -7;
				QB_CU(cudaMemsetAsync(d_ziv.p, 0, 8, stream));
				QB_CU(cudaStreamSynchronize(stream));
			}
			cx.ziv = d_ziv.as<uint32_t>();
			has_ctx = true;
			return 0;
		}
		// -8 (QB_ERR_ROUNDING) once for every new power whose rounding stayed in doubt at the second Ziv level
		int check_rounding() override
		{
			uint64_t c[2];
			if (int rc_ = ziv_counters(c)) return rc_;
			if (c[1] > ziv_reported)
			{
				ziv_reported = c[1];
				return -8;
			}
			return 0;
		}
		int set_ziv_slack(int bits) override
		{
			if (bits < 0) return -4;
			ziv_slack = bits ? bits : QB_ZIV_SLACK;
			cx.ziv_slack = ziv_slack;
			return 0;
		}
		// [0] powers re-evaluated at the second level, [1] powers whose rounding was still in doubt there
		int ziv_counters(uint64_t* out2) override
		{
			out2[0] = out2[1] = 0;
			if (!d_ziv.p) return 0;
			uint32_t c[2] = {0, 0};
			QB_CU(cudaSetDevice(device));
			QB_CU(cudaMemcpyAsync(c, d_ziv.p, 8, cudaMemcpyDeviceToHost, stream));
			QB_CU(cudaStreamSynchronize(stream));
			out2[0] = c[0];
			out2[1] = c[1];
			return 0;
		}
		int context_get(uint32_t* rho, uint32_t* mat) const override
		{
			if (!has_ctx) return -3;
			if (rho) std::memcpy(rho, &consts.rho, sizeof(T));
			if (mat) std::memcpy(mat, consts.mat.data(), sizeof(T) * consts.mat.size());
			return 0;
		}
		int table_dim() const override { return has_ctx ? cf_dim(int(lambda)) : -3; }
		int block_dim(int sym) const override { return has_ctx ? cf_dim_sym(int(lambda), sym) : -3; }

		// b[0] = (-1)^spin (2 eps)_spin / (2^spin (eps)_spin), exact rational rounded once   (src/hor_recursion.cpp:17-19)
		int leading_coefficient(T& out, uint32_t spin)
		{
			auto it = b0_cache.find(spin);
			if (it != b0_cache.end())
			{
				out = it->second;
				return 0;
			}
			constexpr int NB = 160;  // 5120-bit exact integers
			if (spin > 400) return -6;
			mpf<NB> num, den;
			set_si(num, 1, 32 * NB);
			set_si(den, 1, 32 * NB);
			for (uint32_t i = 0; i < spin; ++i)
			{
				int64_t fn = 2 * eps.num + int64_t(i) * int64_t(eps.den);
				int64_t fd = eps.num + int64_t(i) * int64_t(eps.den);
				mul_si(num, num, fn, 32 * NB);
				mul_si(den, den, 2 * fd, 32 * NB);
				if ((kind(num) == K_REG && num.exp > 32 * NB - 64) || (kind(den) == K_REG && den.exp > 32 * NB - 64)) return -6;
			}
			if (is_zero(den)) return -6;
			if (is_zero(num))
				set_zero(out);
			else
				div_wide<L, NB>(out, num, den, prec);
			if (spin & 1u) neg(out);
			b0_cache[spin] = out;
			return 0;
		}
		static bool positive_integer(const T& x) { return kind(x) == K_REG && !sign(x) && is_integer(x); }
		// PrimaryOperator::is_divergent_hor (include/qboot/primary_op.hpp:68-77)
		bool is_divergent(const T& D, uint32_t spin) const
		{
			T t, u;
			si_sub(t, 1 - int64_t(spin), D, prec);
			if (positive_integer(t)) return true;
			mul_si(u, D, 2, prec);
			sub_q(t, u, 2 * int64_t(eps.den) + 2 * eps.num, eps.den, prec);
			neg(t);
			if (positive_integer(t)) return true;
			if (spin > 0)
			{
				sub_q(t, D, (1 + int64_t(spin)) * int64_t(eps.den) + 2 * eps.num, eps.den, prec);
				neg(t);
				if (positive_integer(t)) return true;
			}
			return false;
		}

		int gblock_plan(int n, const uint32_t* spin, const uint32_t* delta, const uint32_t* S, const uint32_t* P) override
		{
			if (!has_ctx) return -3;
			if (n < 0 || (n > 0 && (!spin || !delta || !S || !P))) return -4;
			QB_CU(cudaSetDevice(device));
			// QB_PLAN_TRACE=1: where the planner's host time goes (stderr)
			static const bool plan_trace = std::getenv("QB_PLAN_TRACE") != nullptr;
			auto t_last = std::chrono::steady_clock::now();
			auto lap = [&](const char* what) {
				if (!plan_trace) return;
				const auto t = std::chrono::steady_clock::now();
				std::fprintf(stderr, "[qb plan] %-24s %7.2f ms\n", what, std::chrono::duration<double, std::milli>(t - t_last).count());
				t_last = t;
			};
			const T* hd = reinterpret_cast<const T*>(delta);
			const T* hS = reinterpret_cast<const T*>(S);
			const T* hP = reinterpret_cast<const T*>(P);
			std::vector<T> jd, jS, jP;
			std::vector<uint32_t> jsp;
			std::vector<TableRef> tr(static_cast<size_t>(n));
			T small, onep, onem;
			set_ui_2exp(small, 1, -(3 * prec) / 8 + 15, prec);  // 2^(-(3 prec)/8 + 15)          (hor_recursion.cpp:23)
			add_si(onep, small, 1, prec);
			neg(small);
			add_si(onem, small, 1, prec);
			uint32_t max_spin = 0;
			for (int t = 0; t < n; ++t)
			{
				if ((hd[t].sk & 2u) || (hS[t].sk & 2u) || (hP[t].sk & 2u)) return -4;  // inf / nan
				if (spin[t] > 400) return -6;
				max_spin = std::max(max_spin, spin[t]);
			}
			// b[0] depends on the spin only: a table indexed by spin (401 records at most) instead of a record per job
			std::vector<T> b0tab(size_t(max_spin) + 1);
			{
				std::vector<uint8_t> seen(size_t(max_spin) + 1, 0);
				for (auto& x : b0tab) set_zero(x);
				for (int t = 0; t < n; ++t)
					if (!seen[spin[t]])
					{
						seen[spin[t]] = 1;
						if (int rc = leading_coefficient(b0tab[spin[t]], spin[t])) return rc;
					}
			}
			lap("b[0] by spin");
			// the divergence test is the only per-table arithmetic of the planner: fan it out over host threads
			std::vector<uint8_t> diverges(static_cast<size_t>(n), 0);
			bool any_divergent = false;
			{
				const int nth = n >= 4096 ? int(std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency()))) : 1;
				auto work = [&](int lo, int hi) {
					for (int t = lo; t < hi; ++t) diverges[size_t(t)] = (!is_zero(hd[t]) && is_divergent(hd[t], spin[t])) ? 1 : 0;
				};
				std::vector<std::thread> pool;
				for (int k = 1; k < nth; ++k) pool.emplace_back(work, int(int64_t(n) * k / nth), int(int64_t(n) * (k + 1) / nth));
				work(0, int(int64_t(n) / nth));
				for (auto& th : pool) th.join();
				for (int t = 0; t < n && !any_divergent; ++t) any_divergent = diverges[size_t(t)] != 0;
			}
			lap("divergence tests");
			// Common case, no divergent operator: job t IS table t -- the device-side job arrays alias the table arrays and
			// nothing is repacked (for a scan-sized batch the repacking and its upload were ~0.1 s of host time per call).
			if (!any_divergent)
				for (int t = 0; t < n; ++t) tr[size_t(t)] = TableRef{t, -1};
			else
			{
				jd.reserve(size_t(n));
				jS.reserve(size_t(n));
				jP.reserve(size_t(n));
				jsp.reserve(size_t(n));
				for (int t = 0; t < n; ++t)
				{
					auto push = [&](const T& D) {
						jd.push_back(D);
						jS.push_back(hS[t]);
						jP.push_back(hP[t]);
						jsp.push_back(spin[t]);
						return int32_t(jd.size() - 1);
					};
					if (diverges[size_t(t)])
					{
						// two evaluations at Delta (1 +- s), averaged                         (primary_op.hpp:53-56)
						T dp, dm;
						mul(dp, hd[t], onep, prec);
						mul(dm, hd[t], onem, prec);
						tr[size_t(t)].job0 = push(dp);
						tr[size_t(t)].job1 = push(dm);
					}
					else
					{
						tr[size_t(t)].job0 = push(hd[t]);
						tr[size_t(t)].job1 = -1;
					}
				}
			}
			// from here on the previous plan is gone; the new one counts only once everything below has succeeded
			n_tables = 0;
			tables_valid = false;
			n_jobs = any_divergent ? int(jd.size()) : n;
			job_of_table.assign(size_t(n) + 1, n_jobs);
			for (int t = 0; t < n; ++t) job_of_table[size_t(t)] = tr[size_t(t)].job0;
			// Chunks of CHUNK tables (the last one shorter); the scaled-coefficient scratch is one chunk-sized slot.
			// CHUNK is a whole number of k_horner waves (one thread per (table, k), 4 CTAs of 128 threads per SM), three by
			// default: partial waves leave the wide stream's SMs idle at every kernel boundary (measured: 3 waves 150 k
			// tables/s, 1.6 waves 138 k, 1 wave 128 k at the north-star size).  The scratch is capped at 24 GB.
			{
				static const int env_chunk = std::getenv("QB_CHUNK") ? std::max(256, std::atoi(std::getenv("QB_CHUNK"))) : 0;
				if (sm_count == 0)
				{
					cudaDeviceProp prop;
					QB_CU(cudaGetDeviceProperties(&prop, device));
					sm_count = prop.multiProcessorCount;
				}
				const int64_t wave = std::max<int64_t>(1, int64_t(sm_count) * 4 * 128 / (int64_t(lambda) + 1));
				const int64_t per_table = int64_t(sizeof(T)) * (int64_t(lambda) + 1) * (int64_t(n_max) + 1);
				int64_t chunk = env_chunk ? env_chunk : 3 * wave;
				const int64_t cap = std::max<int64_t>(256, (int64_t(24) << 30) / per_table);
				if (chunk > cap) chunk = cap >= wave ? cap / wave * wave : cap;
				if ((n + chunk - 1) / chunk > 32) chunk = ((int64_t(n) + 31) / 32 + wave - 1) / wave * wave;  // at most 32 chunks
				n_chunks = n > chunk ? int((n + chunk - 1) / chunk) : 1;
				chunk_t.assign(size_t(n_chunks) + 1, n);
				for (int c = 0; c < n_chunks; ++c) chunk_t[size_t(c)] = int(std::min<int64_t>(n, chunk * c));
				// front groups: whole chunks, about three waves of k_series (one thread per job, one CTA of 384 per SM) each
				static const int env_front = std::getenv("QB_FRONT") ? std::max(256, std::atoi(std::getenv("QB_FRONT"))) : 0;
				const int64_t front = env_front ? env_front : int64_t(3) * sm_count * 384;
				group_c.assign(1, 0);
				for (int c = 1; c <= n_chunks; ++c)
					if (c == n_chunks || chunk_t[size_t(c)] - chunk_t[size_t(group_c.back())] >= front) group_c.push_back(c);
				// a short last group joins its predecessor
				if (group_c.size() > 2 && n - chunk_t[size_t(group_c[group_c.size() - 2])] < front / 2) group_c.erase(group_c.end() - 2);
			}
			int max_chunk = 0;
			for (int c = 0; c < n_chunks; ++c) max_chunk = std::max(max_chunk, chunk_t[size_t(c) + 1] - chunk_t[size_t(c)]);
			a_slot_records = (size_t(lambda) + 1) * (size_t(n_max) + 1) * size_t(n_chunks > 1 ? max_chunk : n);
			if (n == 0) return 0;
			lap("jobs and chunks");
			const size_t nt = size_t(n), nj = size_t(n_jobs), dimM = size_t(cf_dim(int(lambda))), n1 = size_t(lambda) + 1;
			if (d_in.ensure(sizeof(T) * 3 * nt + 4 * nt) || (any_divergent && d_jobs.ensure(sizeof(T) * 3 * nj + 4 * nj)) ||
			    d_b0tab.ensure(sizeof(T) * b0tab.size()) ||
			    d_tref.ensure(sizeof(TableRef) * nt) || d_coef.ensure(sizeof(T) * QB_NCOEF * nj) ||
			    d_coefx.ensure(4 * size_t(fx_words<L>()) * QB_NCOEF * (nj + 32))  /* blocks of 32 jobs per launch: the last one may be partial */ ||
			    d_b.ensure(sizeof(T) * (size_t(n_max) + 1) * nj) || d_fr.ensure(sizeof(T) * n1 * nt) || d_pw.ensure(sizeof(T) * (2 * n1 + 1) * nt) || d_qc.ensure(sizeof(T) * nt) || d_a.ensure(sizeof(T) * a_slot_records) ||
			    d_g.ensure(sizeof(T) * dimM * nt))
				return -7;
			lap("device buffers");
			t_delta = d_in.as<T>();
			t_S = t_delta + nt;
			t_P = t_S + nt;
			t_spin = reinterpret_cast<uint32_t*>(t_P + nt);
			QB_CU(cudaMemcpyAsync(t_delta, hd, sizeof(T) * nt, cudaMemcpyHostToDevice, stream));
			QB_CU(cudaMemcpyAsync(t_S, hS, sizeof(T) * nt, cudaMemcpyHostToDevice, stream));
			QB_CU(cudaMemcpyAsync(t_P, hP, sizeof(T) * nt, cudaMemcpyHostToDevice, stream));
			QB_CU(cudaMemcpyAsync(t_spin, spin, 4 * nt, cudaMemcpyHostToDevice, stream));
			if (any_divergent)
			{
				j_delta = d_jobs.as<T>();
				j_S = j_delta + nj;
				j_P = j_S + nj;
				j_spin = reinterpret_cast<uint32_t*>(j_P + nj);
				QB_CU(cudaMemcpyAsync(j_delta, jd.data(), sizeof(T) * nj, cudaMemcpyHostToDevice, stream));
				QB_CU(cudaMemcpyAsync(j_S, jS.data(), sizeof(T) * nj, cudaMemcpyHostToDevice, stream));
				QB_CU(cudaMemcpyAsync(j_P, jP.data(), sizeof(T) * nj, cudaMemcpyHostToDevice, stream));
				QB_CU(cudaMemcpyAsync(j_spin, jsp.data(), 4 * nj, cudaMemcpyHostToDevice, stream));
			}
			else
			{
				j_delta = t_delta;
				j_S = t_S;
				j_P = t_P;
				j_spin = t_spin;
			}
			QB_CU(cudaMemcpyAsync(d_b0tab.p, b0tab.data(), sizeof(T) * b0tab.size(), cudaMemcpyHostToDevice, stream));
			QB_CU(cudaMemcpyAsync(d_tref.p, tr.data(), sizeof(TableRef) * nt, cudaMemcpyHostToDevice, stream));
			QB_CU(cudaStreamSynchronize(stream));  // host vectors go out of scope
			lap("uploads");
			n_tables = n;
			return 0;
		}

		// ---- the stages of one chunk (tables [t0, t1), jobs [j0, j1)) ---------------------------------------------------
		struct Range
		{
			int t0, t1, j0, j1;
		};
		Range range_of(int c) const
		{
			const int t0 = chunk_t[size_t(c)], t1 = chunk_t[size_t(c) + 1];
			return Range{t0, t1, job_of_table[size_t(t0)], job_of_table[size_t(t1)]};
		}
		void stage_rec(const Range& r, cudaStream_t st)
		{
			launch_rec_coeffs<L>(cx, r.j1 - r.j0, j_delta + r.j0, j_spin + r.j0, j_S + r.j0, j_P + r.j0,
			                     d_coef.as<T>() + size_t(r.j0) * QB_NCOEF, st);
			launch_coef_fx<L>(cx, r.j1 - r.j0, j_spin + r.j0, d_coef.as<T>() + size_t(r.j0) * QB_NCOEF,
			                  d_coefx.as<uint32_t>() + size_t(r.j0) * QB_NCOEF * fx_words<L>(), st);
			launches += 2;
		}
		void stage_series(const Range& r, cudaStream_t st)
		{
			launch_series<L>(cx, r.j1 - r.j0, j_delta + r.j0, j_spin + r.j0, d_b0tab.as<T>(), d_coef.as<T>() + size_t(r.j0) * QB_NCOEF,
			                 d_coefx.as<uint32_t>() + size_t(r.j0) * QB_NCOEF * fx_words<L>(), d_b.as<T>() + size_t(r.j0) * (size_t(n_max) + 1), st);
			++launches;
		}
		void stage_pow(const Range& r, cudaStream_t st)
		{
			launch_pow<L>(cx, r.t1 - r.t0, t_delta + r.t0, d_pw.as<T>() + size_t(r.t0) * (2 * size_t(lambda) + 3), st);
			++launches;
		}
		void stage_deriv(const Range& r, cudaStream_t st)
		{
			const size_t nk = size_t(lambda) + 1;
			T* pw = d_pw.as<T>() + size_t(r.t0) * (2 * nk + 1);
			// TableRef holds absolute job indices: hand the whole series array.  One scratch slot: the (scale, horner)
			// pairs of all chunks are serialised on one stream.
			launch_scale<L>(cx, r.t1 - r.t0, d_tref.as<TableRef>() + r.t0, d_b.as<T>(), pw, d_a.as<T>(), st);
			++launches;
			launch_horner<L>(cx, r.t1 - r.t0, d_a.as<T>(), pw, d_fr.as<T>() + size_t(r.t0) * nk, st);
			++launches;
		}
		void stage_finish(const Range& r, cudaStream_t st)
		{
			const size_t nk = size_t(lambda) + 1, dimM = size_t(cf_dim(int(lambda)));
			launches += launch_finish<L>(cx, r.t1 - r.t0, t_delta + r.t0, t_spin + r.t0, t_S + r.t0, t_P + r.t0,
			                             d_fr.as<T>() + size_t(r.t0) * nk, d_g.as<T>() + size_t(r.t0) * dimM, d_qc.as<T>() + r.t0, st);
		}
		// copy of one finished chunk into fetch_dst / sink (host memory, or device memory of this or a peer GPU: the copy
		// engine moves it over PCIe or NVLink underneath the kernels of the later chunks)
		uint32_t* out_dst() const { return fetch_dst ? fetch_dst : sink; }
		int copy_chunk_out(const Range& r, cudaStream_t st)
		{
			const size_t rec = sizeof(T) * size_t(cf_dim(int(lambda)));
			QB_CU(cudaMemcpyAsync(reinterpret_cast<char*>(out_dst()) + rec * size_t(r.t0), d_g.as<char>() + rec * size_t(r.t0),
			                      rec * size_t(r.t1 - r.t0), cudaMemcpyDefault, st));
			return 0;
		}
		// ---- scheduling ---------------------------------------------------------------------------------------------------
		// The batch is cut into chunks (table ranges; the scaled-coefficient scratch is one chunk-sized slot) and the chunks
		// are gathered into FRONT GROUPS of about three waves of the series kernel.  The front stages -- recurrence
		// coefficients, series recurrence, powers: one thread per job / table, a long dependent chain each -- are launched
		// once per group so that every scheduler has its four warps; the wide stages run chunk by chunk:
		//   lo_stream   coefficients + series + powers of every group, in order
		//   mid_stream  scale + Horner of every chunk, in order (they share the one `a` scratch slot)
		//   fin_stream  rho -> z + transverse recursion (instruction-fetch-bound, low issue rate); copy_stream the copy-out
		// so the front stages of group g + 1 run underneath the wide stages of group g and the kernels fill each other's
		// tails.  Events carry the dependencies; everything forks from and joins the handle's stream.
		// With timing enabled (or QB_DEBUG_SYNC / QB_SERIAL) everything runs in order on the handle's stream instead, with an
		// event in front of every stage.
		Range group_range(int g) const
		{
			const int t0 = chunk_t[size_t(group_c[size_t(g)])], t1 = chunk_t[size_t(group_c[size_t(g) + 1])];
			return Range{t0, t1, job_of_table[size_t(t0)], job_of_table[size_t(t1)]};
		}
		// timing: an event in front of stage `stage` (0 coefficients, 1 series, 2 powers + derivatives, 3 finish, -1 end)
		int tick(int stage, cudaStream_t st)
		{
			if (!timing) return 0;
			if (ev_used == ev.size())
			{
				cudaEvent_t x;
				QB_CU(cudaEventCreate(&x));
				ev.push_back(x);
			}
			QB_CU(cudaEventRecord(ev[ev_used++], st));
			ev_stage.push_back(stage);
			return 0;
		}
		// side: the powers go to the side stream (not while timing stages or debugging: one stream then)
		int run_serial(cudaStream_t st, bool side)
		{
			ev_used = 0;
			ev_stage.clear();
			const int n_groups = int(group_c.size()) - 1;
			if (side)
				if (int rc_ = ensure_aux()) return rc_;
			for (int g = 0; g < n_groups; ++g)
			{
				const Range R = group_range(g);
				if (side)
				{
					QB_CU(cudaEventRecord(aux_fork_ev, st));
					QB_CU(cudaStreamWaitEvent(aux_stream, aux_fork_ev, 0));
					stage_pow(R, aux_stream);
					QB_CU(cudaEventRecord(aux_done_ev, aux_stream));
				}
				if (int rc_ = tick(0, st)) return rc_;
				stage_rec(R, st);
				if (int rc_ = debug_sync("k_rec_coeffs")) return rc_;
				if (int rc_ = tick(1, st)) return rc_;
				stage_series(R, st);
				if (int rc_ = debug_sync("k_series")) return rc_;
				if (int rc_ = tick(2, st)) return rc_;
				if (side)
					QB_CU(cudaStreamWaitEvent(st, aux_done_ev, 0));
				else
					stage_pow(R, st);
				if (int rc_ = debug_sync("k_pow")) return rc_;
				for (int c = group_c[size_t(g)]; c < group_c[size_t(g) + 1]; ++c)
				{
					const Range r = range_of(c);
					if (int rc_ = tick(2, st)) return rc_;
					stage_deriv(r, st);
					if (int rc_ = debug_sync("k_scale/k_horner")) return rc_;
					if (int rc_ = tick(3, st)) return rc_;
					stage_finish(r, st);
					if (int rc_ = debug_sync("k_finish")) return rc_;
					if (int rc_ = tick(-1, st)) return rc_;
					if (out_dst())
						if (int rc_ = copy_chunk_out(r, st)) return rc_;
				}
			}
			return tick(-1, st);
		}
		int gblock_run() override
		{
			if (!has_ctx) return -3;
			if (n_tables == 0) return 0;
			QB_CU(cudaSetDevice(device));
			static const bool serial_env = std::getenv("QB_DEBUG_SYNC") != nullptr || std::getenv("QB_SERIAL") != nullptr;
			if (timing || n_chunks == 1 || serial_env)
			{
				if (int rc = run_serial(stream, !timing && !serial_env)) return rc;
				QB_CU(cudaGetLastError());
				tables_valid = true;
				return 0;
			}
			if (!lo_stream)
			{
				int pr_lo = 0, pr_hi = 0;
				QB_CU(cudaDeviceGetStreamPriorityRange(&pr_lo, &pr_hi));
				QB_CU(cudaStreamCreateWithPriority(&lo_stream, cudaStreamNonBlocking, pr_lo));
				QB_CU(cudaStreamCreateWithPriority(&mid_stream, cudaStreamNonBlocking, pr_lo));
				QB_CU(cudaStreamCreateWithPriority(&fin_stream, cudaStreamNonBlocking, pr_lo));
				QB_CU(cudaEventCreateWithFlags(&fork_ev, cudaEventDisableTiming));
				QB_CU(cudaEventCreateWithFlags(&copy_done_ev, cudaEventDisableTiming));
				QB_CU(cudaEventCreateWithFlags(&fin_done_ev, cudaEventDisableTiming));
				QB_CU(cudaEventCreateWithFlags(&lo_done_ev, cudaEventDisableTiming));
				QB_CU(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
			}
			const int n_groups = int(group_c.size()) - 1;
			while (dep_ev.size() < size_t(n_groups) + 2 * size_t(n_chunks))
			{
				cudaEvent_t x;
				QB_CU(cudaEventCreateWithFlags(&x, cudaEventDisableTiming));
				dep_ev.push_back(x);
			}
			auto GE = [&](int g) { return dep_ev[size_t(g)]; };                               // front stages of group g done
			auto EV = [&](int c, int i) { return dep_ev[size_t(n_groups) + size_t(2 * c + i)]; };  // 0 horner done, 1 finish done
			QB_CU(cudaEventRecord(fork_ev, stream));
			for (cudaStream_t s_ : {lo_stream, mid_stream, fin_stream}) QB_CU(cudaStreamWaitEvent(s_, fork_ev, 0));
			for (int g = 0; g < n_groups; ++g)
			{
				const Range R = group_range(g);
				stage_rec(R, lo_stream);
				stage_series(R, lo_stream);
				stage_pow(R, lo_stream);
				QB_CU(cudaEventRecord(GE(g), lo_stream));
			}
			for (int g = 0; g < n_groups; ++g)
			{
				QB_CU(cudaStreamWaitEvent(mid_stream, GE(g), 0));
				for (int c = group_c[size_t(g)]; c < group_c[size_t(g) + 1]; ++c)
				{
					const Range r = range_of(c);
					stage_deriv(r, mid_stream);
					QB_CU(cudaEventRecord(EV(c, 0), mid_stream));
					QB_CU(cudaStreamWaitEvent(fin_stream, EV(c, 0), 0));
					stage_finish(r, fin_stream);
					if (out_dst())
					{
						QB_CU(cudaEventRecord(EV(c, 1), fin_stream));
						QB_CU(cudaStreamWaitEvent(copy_stream, EV(c, 1), 0));
						if (int rc_ = copy_chunk_out(r, copy_stream)) return rc_;
					}
				}
			}
			// join: fin_stream is behind mid_stream, which is behind lo_stream
			QB_CU(cudaEventRecord(fin_done_ev, fin_stream));
			QB_CU(cudaStreamWaitEvent(stream, fin_done_ev, 0));
			QB_CU(cudaEventRecord(lo_done_ev, lo_stream));
			QB_CU(cudaStreamWaitEvent(stream, lo_done_ev, 0));
			if (out_dst())
			{
				QB_CU(cudaEventRecord(copy_done_ev, copy_stream));
				QB_CU(cudaStreamWaitEvent(stream, copy_done_ev, 0));
			}
			QB_CU(cudaGetLastError());
			tables_valid = true;
			return 0;
		}
		// after a synchronise: read the per-stage times of the last run
		int collect_times()
		{
			for (float& x : kernel_ms) x = 0.f;
			if (!timing) return 0;
			for (size_t i = 0; i + 1 < ev_used && i + 1 < ev_stage.size(); ++i)
			{
				if (ev_stage[i] < 0) continue;
				float ms = 0.f;
				if (cudaEventElapsedTime(&ms, ev[i], ev[i + 1]) != cudaSuccess)
				{
					cudaGetLastError();
					ms = 0.f;
				}
				kernel_ms[ev_stage[i]] += ms;
			}
			return 0;
		}
		int collect_times_public() override { return collect_times(); }
		int imad_peak(int iters, double* macs_per_s) override
		{
			QB_CU(cudaSetDevice(device));
			if (d_fb_out.ensure(64)) return -7;
			cudaDeviceProp prop;
			QB_CU(cudaGetDeviceProperties(&prop, device));
			const int blocks = prop.multiProcessorCount * 8, threads = 256;
			cudaEvent_t e0, e1;
			QB_CU(cudaEventCreate(&e0));
			QB_CU(cudaEventCreate(&e1));
			launch_imad_peak(d_fb_out.as<uint64_t>(), 64, 12345u, blocks, threads, stream);  // warm-up
			double best = 0.0;
			for (int rep = 0; rep < 5; ++rep)
			{
				QB_CU(cudaEventRecord(e0, stream));
				launch_imad_peak(d_fb_out.as<uint64_t>(), iters, 12345u + rep, blocks, threads, stream);
				QB_CU(cudaEventRecord(e1, stream));
				QB_CU(cudaStreamSynchronize(stream));
				float ms = 0.f;
				QB_CU(cudaEventElapsedTime(&ms, e0, e1));
				double macs = double(blocks) * threads * double(iters) * 64.0;
				if (ms > 0.f && macs / (ms * 1e-3) > best) best = macs / (ms * 1e-3);
			}
			launches += 6;
			cudaEventDestroy(e0);
			cudaEventDestroy(e1);
			*macs_per_s = best;
			return 0;
		}
		int gblock_fetch(uint32_t* out) override
		{
			if (!has_ctx) return -3;
			if (n_tables == 0) return 0;
			if (!tables_valid) return -4;
			QB_CU(cudaSetDevice(device));
			size_t bytes = sizeof(T) * size_t(cf_dim(int(lambda))) * size_t(n_tables);
			QB_CU(cudaMemcpyAsync(out, d_g.p, bytes, cudaMemcpyDeviceToHost, stream));
			QB_CU(cudaStreamSynchronize(stream));
			return check_rounding();
		}
		const uint32_t* device_tables() const override { return d_g.as<uint32_t>(); }

		int fblock_batch(int m, const int32_t* table_idx, int n_d, const uint32_t* d, const int32_t* d_idx,
		                 const int32_t* sym, uint32_t* out) override
		{
			if (!has_ctx) return -3;
			if (m < 0 || n_d < 0) return -4;
			if (m == 0) return 0;
			if (!table_idx || !d || !d_idx || !sym || !out || n_d == 0 || !tables_valid) return -4;
			QB_CU(cudaSetDevice(device));
			const int lam = int(lambda), dimM = cf_dim(lam);
			const int dim_max = cf_dim_sym(lam, 0) > cf_dim_sym(lam, 1) ? cf_dim_sym(lam, 0) : cf_dim_sym(lam, 1);
			std::vector<TermRef> terms(static_cast<size_t>(m));
			int64_t ofs = 0;
			for (int i = 0; i < m; ++i)
			{
				if (table_idx[i] < 0 || table_idx[i] >= n_tables || d_idx[i] < 0 || d_idx[i] >= n_d || (sym[i] != 0 && sym[i] != 1))
					return -4;
				terms[size_t(i)] = TermRef{table_idx[i], d_idx[i], sym[i], int32_t(ofs)};
				ofs += cf_dim_sym(lam, sym[i]);
				if (ofs > 0x7fffffff) return -4;
			}
			if (d_fb_in.ensure(sizeof(T) * size_t(n_d)) || d_fb_v.ensure(sizeof(T) * size_t(n_d) * size_t(dimM)) ||
			    d_fb_out.ensure(sizeof(T) * size_t(ofs)) || d_fb_terms.ensure(sizeof(TermRef) * size_t(m)))
				return -7;
			QB_CU(cudaMemcpyAsync(d_fb_in.p, d, sizeof(T) * size_t(n_d), cudaMemcpyHostToDevice, stream));
			QB_CU(cudaMemcpyAsync(d_fb_terms.p, terms.data(), sizeof(TermRef) * size_t(m), cudaMemcpyHostToDevice, stream));
			{
				launch_vtod<L>(cx, n_d, d_fb_in.as<T>(), d_fb_v.as<T>(), stream);
				++launches;
				if (int rc_ = debug_sync("k_vtod")) return rc_;
			}
			{
				launch_fblock<L>(cx, m, dim_max, d_fb_terms.as<TermRef>(), d_g.as<T>(), d_fb_v.as<T>(), d_fb_out.as<T>(), stream);
				++launches;
				if (int rc_ = debug_sync("k_fblock")) return rc_;
			}
			QB_CU(cudaGetLastError());
			QB_CU(cudaMemcpyAsync(out, d_fb_out.p, sizeof(T) * size_t(ofs), cudaMemcpyDeviceToHost, stream));
			QB_CU(cudaStreamSynchronize(stream));
			return 0;
		}
		// v^d tables of n_d values (HOST in, HOST out: n_d * dim_M records)                     (src/context.cpp:28-42)
		int vtod_batch(int n_d, const uint32_t* d, uint32_t* out) override
		{
			if (!has_ctx) return -3;
			if (n_d <= 0) return n_d == 0 ? 0 : -4;
			if (!d || !out) return -4;
			QB_CU(cudaSetDevice(device));
			const size_t dimM = size_t(cf_dim(int(lambda)));
			if (d_fb_in.ensure(sizeof(T) * size_t(n_d)) || d_fb_v.ensure(sizeof(T) * size_t(n_d) * dimM)) return -7;
			QB_CU(cudaMemcpyAsync(d_fb_in.p, d, sizeof(T) * size_t(n_d), cudaMemcpyHostToDevice, stream));
			launch_vtod<L>(cx, n_d, d_fb_in.as<T>(), d_fb_v.as<T>(), stream);
			++launches;
			QB_CU(cudaGetLastError());
			QB_CU(cudaMemcpyAsync(out, d_fb_v.p, sizeof(T) * size_t(n_d) * dimM, cudaMemcpyDeviceToHost, stream));
			QB_CU(cudaStreamSynchronize(stream));
			return 0;
		}
		// F/H blocks of m (table, d, sym) triples with the TABLES GIVEN BY THE CALLER (HOST, m * dim_M records): term i uses
		// table i and d[i].  Leaves the resident tables of the last gblock_run untouched.
		int fblock_tables(int m, const uint32_t* g, const uint32_t* d, const int32_t* sym, uint32_t* out) override
		{
			if (!has_ctx) return -3;
			if (m <= 0) return m == 0 ? 0 : -4;
			if (!g || !d || !sym || !out) return -4;
			QB_CU(cudaSetDevice(device));
			const int lam = int(lambda), dimM = cf_dim(lam);
			const int dim_max = std::max(cf_dim_sym(lam, 0), cf_dim_sym(lam, 1));
			std::vector<TermRef> terms(static_cast<size_t>(m));
			int64_t ofs = 0;
			for (int i = 0; i < m; ++i)
			{
				if (sym[i] != 0 && sym[i] != 1) return -4;
				terms[size_t(i)] = TermRef{i, i, sym[i], int32_t(ofs)};
				ofs += cf_dim_sym(lam, sym[i]);
				if (ofs > 0x7fffffff) return -4;
			}
			if (d_fb_in.ensure(sizeof(T) * size_t(m)) || d_fb_v.ensure(sizeof(T) * size_t(m) * size_t(dimM)) || d_fb_g.ensure(sizeof(T) * size_t(m) * size_t(dimM)) ||
			    d_fb_out.ensure(sizeof(T) * size_t(ofs)) || d_fb_terms.ensure(sizeof(TermRef) * size_t(m)))
				return -7;
			QB_CU(cudaMemcpyAsync(d_fb_in.p, d, sizeof(T) * size_t(m), cudaMemcpyHostToDevice, stream));
			QB_CU(cudaMemcpyAsync(d_fb_g.p, g, sizeof(T) * size_t(m) * size_t(dimM), cudaMemcpyHostToDevice, stream));
			QB_CU(cudaMemcpyAsync(d_fb_terms.p, terms.data(), sizeof(TermRef) * size_t(m), cudaMemcpyHostToDevice, stream));
			launch_vtod<L>(cx, m, d_fb_in.as<T>(), d_fb_v.as<T>(), stream);
			launch_fblock<L>(cx, m, dim_max, d_fb_terms.as<TermRef>(), d_fb_g.as<T>(), d_fb_v.as<T>(), d_fb_out.as<T>(), stream);
			launches += 2;
			QB_CU(cudaGetLastError());
			QB_CU(cudaMemcpyAsync(out, d_fb_out.p, sizeof(T) * size_t(ofs), cudaMemcpyDeviceToHost, stream));
			QB_CU(cudaStreamSynchronize(stream));
			return 0;
		}
		int op_batch(int op, int n, const uint32_t* a, const uint32_t* b, const uint32_t* c, const int64_t* ia,
		             const int64_t* ib, uint32_t* out) override
		{
			if (!has_ctx) return -3;
			if (n <= 0 || !out) return n == 0 ? 0 : -4;
			QB_CU(cudaSetDevice(device));
			const size_t rec = sizeof(T) * size_t(n), ints = 8 * size_t(n);
			if (d_fb_in.ensure(4 * rec + 2 * ints)) return -7;
			uint8_t* base = d_fb_in.as<uint8_t>();
			T* da = reinterpret_cast<T*>(base);
			T* db = reinterpret_cast<T*>(base + rec);
			T* dc = reinterpret_cast<T*>(base + 2 * rec);
			T* dout = reinterpret_cast<T*>(base + 3 * rec);
			int64_t* dia = reinterpret_cast<int64_t*>(base + 4 * rec);
			int64_t* dib = reinterpret_cast<int64_t*>(base + 4 * rec + ints);
			if (a) QB_CU(cudaMemcpyAsync(da, a, rec, cudaMemcpyHostToDevice, stream));
			if (b) QB_CU(cudaMemcpyAsync(db, b, rec, cudaMemcpyHostToDevice, stream));
			if (c) QB_CU(cudaMemcpyAsync(dc, c, rec, cudaMemcpyHostToDevice, stream));
			if (ia) QB_CU(cudaMemcpyAsync(dia, ia, ints, cudaMemcpyHostToDevice, stream));
			if (ib) QB_CU(cudaMemcpyAsync(dib, ib, ints, cudaMemcpyHostToDevice, stream));
			launch_ops<L>(cx, op, n, a ? da : nullptr, b ? db : nullptr, c ? dc : nullptr, ia ? dia : nullptr, ib ? dib : nullptr, dout, stream);
			++launches;
			QB_CU(cudaGetLastError());
			QB_CU(cudaMemcpyAsync(out, dout, rec, cudaMemcpyDeviceToHost, stream));
			QB_CU(cudaStreamSynchronize(stream));
			return 0;
		}
#include "engine_pmp.inc"
#undef QB_CU
	};

	template <int L>
	EngineBase* make_engine(int device, int prec)
	{
		auto* e = new (std::nothrow) Engine<L>(device, prec);
		if (!e) return nullptr;
		if (cudaSetDevice(device) != cudaSuccess)
		{
			cudaGetLastError();
			delete e;
			return nullptr;
		}
		if (const char* ss = std::getenv("QB_STACK_BYTES")) cudaDeviceSetLimit(cudaLimitStackSize, size_t(std::atol(ss)));
		if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess)
		{
			cudaGetLastError();
			delete e;
			return nullptr;
		}
		return e;
	}
}  // namespace qb
#endif  // QB_L
