// TEST HARNESS ONLY -- a host stand-in for libqboot_b200.so so that the C++ façade (planner, PMP assembly, writer) can
// be exercised end to end by the CPU test-suite, where there is no GPU.  It implements the part of the C ABI of
// include/qboot_b200.h that the façade calls:
//   * block tables come from the ORACLE (oracle/_ref/libqboot_ref.so, the unmodified reference; dlopen'ed) -- the CUDA
//     tables are pinned against the very same library on the GPU box (tests/test_gpu.py), so the façade sees identical bits;
//   * everything after the tables runs the engine's own `__host__ __device__` building blocks (csrc/pmp.cuh, block.cuh)
//     serially: the same source the CUDA kernels are compiled from.
// Nothing in the product links, loads or ships this file; the product library fails with QB_ERR_NO_DEVICE without a GPU.
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/qboot_b200.h"
#include "../../qboot_b200/csrc/hostmath.hpp"
#include "../../qboot_b200/csrc/pmp.cuh"
#include "qboot/b200/hostops.hpp"

namespace
{
	using namespace qb;
	typedef int (*ref_gblock_fn)(int, uint32_t, uint32_t, const char*, int, const uint32_t*, const uint32_t*, const uint32_t*, const uint32_t*, uint32_t*);
	ref_gblock_fn load_ref()
	{
		static ref_gblock_fn fn = [] {
			const char* path = std::getenv("QB_EMU_REF");
			void* so = dlopen(path ? path : "oracle/_ref/libqboot_ref.so", RTLD_NOW | RTLD_LOCAL | RTLD_DEEPBIND);
			if (!so) return ref_gblock_fn(nullptr);
			return reinterpret_cast<ref_gblock_fn>(dlsym(so, "ref_gblock"));
		}();
		return fn;
	}
	struct EmuBase
	{
		int prec = 0, limbs = 0;
		virtual ~EmuBase() {}
		virtual int context_init(uint32_t, uint32_t, int64_t, uint32_t) = 0;
		virtual int context_get(uint32_t*, uint32_t*) = 0;
		virtual int table_dim() = 0;
		virtual int block_dim(int) = 0;
		virtual int plan(int, const uint32_t*, const uint32_t*, const uint32_t*, const uint32_t*) = 0;
		virtual int run() = 0;
		virtual int fetch(uint32_t*) = 0;
		virtual int vtod(int, const uint32_t*, uint32_t*) = 0;
		virtual int fblock_tables(int, const uint32_t*, const uint32_t*, const int32_t*, uint32_t*) = 0;
		virtual int assemble(int, const int32_t*, const int32_t*, int, const qb_pmp_block*, int, const qb_pmp_term*, int, const int32_t*, int,
		                     const uint32_t*, int, const uint32_t*) = 0;
		virtual int block_add(int, int, const uint32_t*, const uint32_t*) = 0;
		virtual int block_fetch(int, uint32_t*) = 0;
		virtual int pack(int, const int32_t*, int, const int32_t*, int, const int32_t*, const uint32_t*, const uint32_t*) = 0;
		virtual int packed_fetch(int, uint32_t*, uint32_t*) = 0;
		virtual int text(int, int64_t*) = 0;
		virtual int text_declined(int, int64_t*, uint32_t*) = 0;
		virtual int text_patch(int, const int64_t*, const char*, int, const int32_t*, int64_t*) = 0;
		virtual int text_fetch(char*) = 0;
		virtual int info(int64_t*) = 0;
		virtual int dec_format(int, int, const uint32_t*, char*, int, int32_t*) = 0;
		virtual int poly_eval(int, const int32_t*, const uint32_t*, int, const uint32_t*, int, const uint32_t*, int, const qb_eval_item*, uint32_t*) = 0;
	};
	template <int L>
	struct Emu : EmuBase
	{
		using T = mpf<L>;
		bool has_ctx = false;
		uint32_t n_max = 0, lambda = 0;
		Rational eps{0, 1};
		ContextConsts<L> consts;
		std::vector<uint32_t> k_spin;
		std::vector<T> k_delta, k_S, k_P, g;
		bool tables_valid = false;
		std::vector<PmpBlock> blocks, sel;
		std::vector<T> M, Tg, B, C;
		int N = 0, n_free = 0;
		std::vector<uint32_t> p13;
		std::vector<uint32_t> ten_n, ten_n1;
		DecTables tb{};
		std::vector<std::string> lines;  // text of every packed number (B then c)
		std::vector<int32_t> len;
		std::string out_text;
		std::vector<int64_t> sizes;

		int context_init(uint32_t nm, uint32_t lam, int64_t en, uint32_t ed) override
		{
			if (ed == 0 || nm == 0) return QB_ERR_BAD_ARG;
			n_max = nm;
			lambda = lam;
			eps = Rational{en, ed};
			make_context_consts<L>(consts, lam, prec);
			has_ctx = true;
			tables_valid = false;
			blocks.clear();
			sel.clear();
			N = 0;
			return 0;
		}
		int context_get(uint32_t* rho, uint32_t* mat) override
		{
			if (!has_ctx) return QB_ERR_NO_CONTEXT;
			if (rho) std::memcpy(rho, &consts.rho, sizeof(T));
			if (mat) std::memcpy(mat, consts.mat.data(), sizeof(T) * consts.mat.size());
			return 0;
		}
		int table_dim() override { return has_ctx ? cf_dim(int(lambda)) : QB_ERR_NO_CONTEXT; }
		int block_dim(int sym) override { return has_ctx ? cf_dim_sym(int(lambda), sym) : QB_ERR_NO_CONTEXT; }
		int plan(int n, const uint32_t* spin, const uint32_t* delta, const uint32_t* S, const uint32_t* P) override
		{
			if (!has_ctx) return QB_ERR_NO_CONTEXT;
			k_spin.assign(spin, spin + n);
			k_delta.assign(reinterpret_cast<const T*>(delta), reinterpret_cast<const T*>(delta) + n);
			k_S.assign(reinterpret_cast<const T*>(S), reinterpret_cast<const T*>(S) + n);
			k_P.assign(reinterpret_cast<const T*>(P), reinterpret_cast<const T*>(P) + n);
			tables_valid = false;
			return 0;
		}
		int run() override
		{
			ref_gblock_fn ref = load_ref();
			if (!ref) return QB_ERR_NO_DEVICE;
			const int n = int(k_spin.size()), dimM = cf_dim(int(lambda));
			g.resize(size_t(n) * size_t(dimM));
			char dim[64];
			// dim = 2 eps + 2 as an exact rational string
			std::snprintf(dim, sizeof(dim), "%lld/%u", static_cast<long long>(2 * eps.num + 2 * int64_t(eps.den)), eps.den);
			if (n)
			{
				int r = ref(prec, n_max, lambda, dim, n, reinterpret_cast<const uint32_t*>(k_delta.data()), k_spin.data(),
				            reinterpret_cast<const uint32_t*>(k_S.data()), reinterpret_cast<const uint32_t*>(k_P.data()),
				            reinterpret_cast<uint32_t*>(g.data()));
				if (r != dimM) return QB_ERR_CUDA;
			}
			tables_valid = true;
			return 0;
		}
		int fetch(uint32_t* out) override
		{
			if (!tables_valid) return QB_ERR_BAD_ARG;
			std::memcpy(out, g.data(), sizeof(T) * g.size());
			return 0;
		}
		void vtable(T* v, const T& d)
		{
			// 4^-d correctly rounded (the device evaluates exp(-d ln 4) at 96 guard bits: same value)
			T md, p4;
			copy(md, d);
			neg(md);
			qb::host::h_ui_pow(reinterpret_cast<uint32_t*>(&p4), 4, reinterpret_cast<const uint32_t*>(&md), prec);
			v_to_d_table(v, 1, int(lambda), d, p4, prec);
		}
		int vtod(int n_d, const uint32_t* d, uint32_t* out) override
		{
			if (!has_ctx) return QB_ERR_NO_CONTEXT;
			const int dimM = cf_dim(int(lambda));
			for (int i = 0; i < n_d; ++i) vtable(reinterpret_cast<T*>(out) + size_t(i) * dimM, reinterpret_cast<const T*>(d)[i]);
			return 0;
		}
		int fblock_tables(int m, const uint32_t* gt, const uint32_t* d, const int32_t* sym, uint32_t* out) override
		{
			if (!has_ctx) return QB_ERR_NO_CONTEXT;
			const int lam = int(lambda), dimM = cf_dim(lam);
			std::vector<T> v(static_cast<size_t>(dimM));
			T* o = reinterpret_cast<T*>(out);
			for (int i = 0; i < m; ++i)
			{
				vtable(v.data(), reinterpret_cast<const T*>(d)[i]);
				const int dim = cf_dim_sym(lam, sym[i]);
				for (int e = 0; e < dim; ++e)
				{
					int Mx, Nx;
					cf_decode_sym(lam, sym[i], e, Mx, Nx);
					fblock_entry(o[e], v.data(), 1, reinterpret_cast<const T*>(gt) + size_t(i) * dimM, 1, lam, Mx, Nx, prec);
				}
				o += dim;
			}
			return 0;
		}
		int assemble(int n_eq, const int32_t* eq_dim, const int32_t* eq_sym, int n_blocks, const qb_pmp_block* bl, int n_terms,
		             const qb_pmp_term* terms, int n_tab, const int32_t* tab, int n_d, const uint32_t* d, int n_coef,
		             const uint32_t* coef) override
		{
			if (!has_ctx) return QB_ERR_NO_CONTEXT;
			if (n_tab && !tables_valid) return QB_ERR_BAD_ARG;
			const int lam = int(lambda), dimM = cf_dim(lam);
			std::vector<int32_t> eq_ofs(size_t(n_eq) + 1, 0);
			for (int e = 0; e < n_eq; ++e)
			{
				if (n_terms > 0 && eq_dim[e] != cf_dim_sym(lam, eq_sym[e])) return QB_ERR_BAD_ARG;
				eq_ofs[size_t(e) + 1] = eq_ofs[size_t(e)] + eq_dim[e];
			}
			N = eq_ofs[size_t(n_eq)];
			std::vector<T> v(size_t(n_d) * size_t(dimM));
			for (int i = 0; i < n_d; ++i) vtable(v.data() + size_t(i) * dimM, reinterpret_cast<const T*>(d)[i]);
			blocks.assign(size_t(n_blocks), PmpBlock{});
			int64_t rec = 0;
			for (int b = 0; b < n_blocks; ++b)
			{
				PmpBlock& k = blocks[size_t(b)];
				k.n_samples = bl[b].n_samples;
				k.sz = bl[b].sz;
				k.term0 = bl[b].term0;
				k.n_terms = bl[b].n_terms;
				k.ofs = rec;
				k.ofs_t = -1;
				rec += int64_t(pmp_schur(k)) * N;
			}
			M.assign(size_t(rec), T{});
			Tg.clear();
			const PmpTerm* tm = reinterpret_cast<const PmpTerm*>(terms);
			for (const PmpBlock& k : blocks)
			{
				const int nrc = k.sz * (k.sz + 1) / 2;
				for (int rc = 0; rc < nrc; ++rc)
					for (int kk = 0; kk < k.n_samples; ++kk)
						for (int e = 0; e < n_eq; ++e)
							for (int i = 0; i < eq_dim[e]; ++i)
								pmp_entry<L>(M[size_t(k.ofs + (int64_t(rc) * k.n_samples + kk) * N + eq_ofs[size_t(e)] + i)], tm + k.term0, k.n_terms, e, rc, kk,
								             eq_sym[e], i, tab, g.data(), v.data(), reinterpret_cast<const T*>(coef), lam, prec);
			}
			sel.clear();
			(void)n_coef;
			(void)n_tab;
			return 0;
		}
		int block_add(int ns, int sz, const uint32_t* mat, const uint32_t* target) override
		{
			if (N <= 0) return QB_ERR_BAD_ARG;
			PmpBlock k{};
			k.n_samples = ns;
			k.sz = sz;
			k.ofs = int64_t(M.size());
			k.ofs_t = -1;
			const size_t schur = size_t(pmp_schur(k));
			M.insert(M.end(), reinterpret_cast<const T*>(mat), reinterpret_cast<const T*>(mat) + schur * size_t(N));
			if (target)
			{
				k.ofs_t = int64_t(Tg.size());
				Tg.insert(Tg.end(), reinterpret_cast<const T*>(target), reinterpret_cast<const T*>(target) + schur);
			}
			blocks.push_back(k);
			return int(blocks.size()) - 1;
		}
		int block_fetch(int b, uint32_t* out) override
		{
			if (b < 0 || b >= int(blocks.size())) return QB_ERR_BAD_ARG;
			const PmpBlock& k = blocks[size_t(b)];
			std::memcpy(out, M.data() + k.ofs, sizeof(T) * size_t(pmp_schur(k)) * size_t(N));
			return 0;
		}
		int pack(int n_sel, const int32_t* s, int n_elim, const int32_t* lead, int nf, const int32_t* free_idx, const uint32_t* eqn,
		         const uint32_t* eqt) override
		{
			if (nf + n_elim != N) return QB_ERR_BAD_ARG;
			sel.clear();
			int64_t rb = 0, rc = 0;
			for (int j = 0; j < n_sel; ++j)
			{
				if (s[j] < 0 || s[j] >= int(blocks.size())) return QB_ERR_BAD_ARG;
				PmpBlock k = blocks[size_t(s[j])];
				k.ofs_b = rb;
				k.ofs_c = rc;
				rb += int64_t(pmp_schur(k)) * nf;
				rc += pmp_schur(k);
				sel.push_back(k);
			}
			B.assign(size_t(rb), T{});
			C.assign(size_t(rc), T{});
			n_free = nf;
			for (const PmpBlock& k : sel)
				for (int64_t p = 0; p < pmp_schur(k); ++p)
					for (int m = 0; m <= nf; ++m)
					{
						T r;
						pmp_pack_entry<L>(r, M.data() + k.ofs + p * N, m, nf, free_idx, n_elim, lead, reinterpret_cast<const T*>(eqn),
						                  reinterpret_cast<const T*>(eqt), k.ofs_t >= 0 ? Tg.data() + k.ofs_t + p : nullptr, N, prec);
						if (m < nf)
							B[size_t(k.ofs_b + p * nf + m)] = r;
						else
							C[size_t(k.ofs_c + p)] = r;
					}
			return 0;
		}
		int packed_fetch(int j, uint32_t* Bo, uint32_t* co) override
		{
			if (j < 0 || j >= int(sel.size())) return QB_ERR_BAD_ARG;
			const PmpBlock& k = sel[size_t(j)];
			if (Bo) std::memcpy(Bo, B.data() + k.ofs_b, sizeof(T) * size_t(pmp_schur(k)) * size_t(n_free));
			if (co) std::memcpy(co, C.data() + k.ofs_c, sizeof(T) * size_t(pmp_schur(k)));
			return 0;
		}
		int prepare(int ndigits)
		{
			if (ndigits < 1 || double(ndigits) * 3.3219281 > 32.0 * (L + 2) - 2) return QB_ERR_BAD_ARG;
			auto mul_small = [](std::vector<uint32_t>& v, uint32_t m) {
				uint64_t c = 0;
				for (auto& w : v)
				{
					uint64_t t = uint64_t(w) * m + c;
					w = uint32_t(t);
					c = t >> 32;
				}
				if (c) v.push_back(uint32_t(c));
			};
			p13.clear();
			std::vector<uint32_t> cur{1u};
			for (int a = 0; a < QB_DEC_MAXA; ++a)
			{
				tb.pofs[a] = int32_t(p13.size());
				p13.insert(p13.end(), cur.begin(), cur.end());
				mul_small(cur, 1220703125u);
			}
			tb.pofs[QB_DEC_MAXA] = int32_t(p13.size());
			auto pow10 = [&](int k) {
				std::vector<uint32_t> v{1u};
				for (int i = 0; i < k; ++i) mul_small(v, 10u);
				v.resize(size_t(L) + 2, 0u);
				return v;
			};
			ten_n = pow10(ndigits);
			ten_n1 = pow10(ndigits - 1);
			tb.p13 = p13.data();
			tb.ten_n = ten_n.data();
			tb.ten_n1 = ten_n1.data();
			tb.ndigits = ndigits;
			return 0;
		}
		int poly_eval(int n_polys, const int32_t* ofs, const uint32_t* coef, int n_x, const uint32_t* x, int n_s, const uint32_t* s, int n_items,
		              const qb_eval_item* items, uint32_t* out) override
		{
			(void)n_polys;
			(void)n_x;
			(void)n_s;
			const T* c = reinterpret_cast<const T*>(coef);
			for (int i = 0; i < n_items; ++i)
			{
				const int o = ofs[items[i].poly], deg = ofs[items[i].poly + 1] - o - 1;
				poly_eval_scaled<L>(reinterpret_cast<T*>(out)[i], c + o, deg, reinterpret_cast<const T*>(x)[items[i].x],
				                    reinterpret_cast<const T*>(s)[items[i].s], prec);
			}
			return 0;
		}
		int dec_format(int ndigits, int n, const uint32_t* x, char* slots, int slot, int32_t* l) override
		{
			if (int rc = prepare(ndigits)) return rc;
			for (int i = 0; i < n; ++i) l[i] = format_decimal<L>(slots + size_t(i) * size_t(slot), reinterpret_cast<const T*>(x)[i], tb);
			return 0;
		}
		int finish_text(int64_t* sz)
		{
			out_text.clear();
			sizes.assign(2 * sel.size(), 0);
			const int64_t nb = int64_t(B.size());
			for (size_t j = 0; j < sel.size(); ++j)
			{
				const PmpBlock& k = sel[j];
				size_t start = out_text.size();
				for (int64_t i = k.ofs_b; i < k.ofs_b + int64_t(pmp_schur(k)) * n_free; ++i) out_text += lines[size_t(i)];
				sizes[2 * j] = int64_t(out_text.size() - start);
				start = out_text.size();
				for (int64_t i = nb + k.ofs_c; i < nb + k.ofs_c + pmp_schur(k); ++i) out_text += lines[size_t(i)];
				sizes[2 * j + 1] = int64_t(out_text.size() - start);
			}
			for (size_t j = 0; j < sizes.size(); ++j) sz[j] = sizes[j];
			return 0;
		}
		int text(int ndigits, int64_t* sz) override
		{
			if (int rc = prepare(ndigits)) return rc;
			const size_t n = B.size() + C.size();
			lines.assign(n, std::string());
			len.assign(n, 0);
			std::vector<char> slot(size_t(ndigits) + 16);
			int declined = 0;
			for (size_t i = 0; i < n; ++i)
			{
				const T& x = i < B.size() ? B[i] : C[i - B.size()];
				len[i] = format_decimal<L>(slot.data(), x, tb);
				if (len[i] < 0)
					++declined;
				else
					lines[i].assign(slot.data(), size_t(len[i]));
			}
			if (declined) return declined;
			return finish_text(sz);
		}
		int text_declined(int cap, int64_t* idx, uint32_t* recs) override
		{
			int cnt = 0;
			for (size_t i = 0; i < len.size() && cnt < cap; ++i)
				if (len[i] < 0)
				{
					idx[cnt] = int64_t(i);
					const T& x = i < B.size() ? B[i] : C[i - B.size()];
					std::memcpy(recs + size_t(cnt) * (L + 2), &x, sizeof(T));
					++cnt;
				}
			return cnt;
		}
		int text_patch(int cnt, const int64_t* idx, const char* texts, int stride, const int32_t* l, int64_t* sz) override
		{
			for (int i = 0; i < cnt; ++i)
			{
				lines[size_t(idx[i])].assign(texts + size_t(i) * size_t(stride), size_t(l[i]));
				len[size_t(idx[i])] = l[i];
			}
			return finish_text(sz);
		}
		int text_fetch(char* out) override
		{
			std::memcpy(out, out_text.data(), out_text.size());
			return 0;
		}
		int info(int64_t* o) override
		{
			o[0] = int64_t(blocks.size());
			o[1] = N;
			o[2] = int64_t(M.size());
			o[3] = int64_t(sel.size());
			return 0;
		}
	};
}  // namespace

struct qb_handle
{
	EmuBase* e;
};

extern "C"
{
	const char* qb_version(void) { return "qboot_b200 host stand-in (tests only)"; }
	const char* qb_strerror(int s) { return s == 0 ? "ok" : s == QB_ERR_NO_DEVICE ? "oracle library not found (emu)" : "error (emu)"; }
	int qb_precision_supported(int prec) { return qb::host::limbs_for(prec) != 0; }
	int qb_create(int, int prec, qb_handle** out)
	{
		EmuBase* e = nullptr;
		switch ((prec + 31) / 32)
		{
#define X(LL) \
	case LL: e = new Emu<LL>(); break;
			X(16) X(19) X(25) X(32) X(38) X(48)
#undef X
		    default : return QB_ERR_BAD_PREC;
		}
		e->prec = prec;
		e->limbs = (prec + 31) / 32;
		*out = new qb_handle{e};
		return 0;
	}
	void qb_destroy(qb_handle* h)
	{
		if (!h) return;
		delete h->e;
		delete h;
	}
	int qb_words(const qb_handle* h) { return h->e->limbs + 2; }
	const char* qb_last_cuda_error(const qb_handle*) { return ""; }
	int qb_context_init(qb_handle* h, uint32_t n_max, uint32_t lambda, int64_t en, uint32_t ed) { return h->e->context_init(n_max, lambda, en, ed); }
	int qb_context_get(const qb_handle* h, uint32_t* rho, uint32_t* mat) { return h->e->context_get(rho, mat); }
	int qb_table_dim(const qb_handle* h) { return h->e->table_dim(); }
	int qb_block_dim(const qb_handle* h, int sym) { return h->e->block_dim(sym); }
	int qb_gblock_plan(qb_handle* h, int n, const uint32_t* spin, const uint32_t* delta, const uint32_t* S, const uint32_t* P)
	{
		return h->e->plan(n, spin, delta, S, P);
	}
	int qb_gblock_run(qb_handle* h) { return h->e->run(); }
	int qb_gblock_fetch(qb_handle* h, uint32_t* out) { return h->e->fetch(out); }
	int qb_gblock_batch(qb_handle* h, int n, const uint32_t* spin, const uint32_t* delta, const uint32_t* S, const uint32_t* P, uint32_t* out)
	{
		int rc = h->e->plan(n, spin, delta, S, P);
		if (!rc) rc = h->e->run();
		if (!rc) rc = h->e->fetch(out);
		return rc;
	}
	int qb_sync(qb_handle*) { return 0; }
	int qb_vtod_batch(qb_handle* h, int n_d, const uint32_t* d, uint32_t* out) { return h->e->vtod(n_d, d, out); }
	int qb_fblock_tables(qb_handle* h, int m, const uint32_t* g, const uint32_t* d, const int32_t* sym, uint32_t* out)
	{
		return h->e->fblock_tables(m, g, d, sym, out);
	}
	void* qb_pinned_alloc(size_t bytes) { return bytes ? std::malloc(bytes) : nullptr; }
	void qb_pinned_free(void* p) { std::free(p); }
	int qb_pmp_assemble(qb_handle* h, int n_eq, const int32_t* eq_dim, const int32_t* eq_sym, int n_blocks, const qb_pmp_block* blocks, int n_terms,
	                    const qb_pmp_term* terms, int n_tab, const int32_t* tab, int n_d, const uint32_t* d, int n_coef, const uint32_t* coef)
	{
		return h->e->assemble(n_eq, eq_dim, eq_sym, n_blocks, blocks, n_terms, terms, n_tab, tab, n_d, d, n_coef, coef);
	}
	int qb_pmp_block_add(qb_handle* h, int ns, int sz, const uint32_t* mat, const uint32_t* target) { return h->e->block_add(ns, sz, mat, target); }
	int qb_pmp_block_fetch(qb_handle* h, int b, uint32_t* out) { return h->e->block_fetch(b, out); }
	int qb_pmp_pack(qb_handle* h, int n_sel, const int32_t* sel, int n_elim, const int32_t* lead, int n_free, const int32_t* free_idx,
	                const uint32_t* eqn, const uint32_t* eqt)
	{
		return h->e->pack(n_sel, sel, n_elim, lead, n_free, free_idx, eqn, eqt);
	}
	int qb_pmp_packed_fetch(qb_handle* h, int j, uint32_t* B, uint32_t* c) { return h->e->packed_fetch(j, B, c); }
	int qb_pmp_text(qb_handle* h, int ndigits, int64_t* sizes) { return h->e->text(ndigits, sizes); }
	int qb_pmp_text_declined(qb_handle* h, int cap, int64_t* idx, uint32_t* recs) { return h->e->text_declined(cap, idx, recs); }
	int qb_pmp_text_patch(qb_handle* h, int cnt, const int64_t* idx, const char* texts, int stride, const int32_t* len, int64_t* sizes)
	{
		return h->e->text_patch(cnt, idx, texts, stride, len, sizes);
	}
	int qb_pmp_text_fetch(qb_handle* h, char* out) { return h->e->text_fetch(out); }
	int qb_pmp_info(const qb_handle* h, int64_t* out4) { return h->e->info(out4); }
	int qb_dec_format(qb_handle* h, int ndigits, int n, const uint32_t* x, char* slots, int slot, int32_t* len)
	{
		return h->e->dec_format(ndigits, n, x, slots, slot, len);
	}
	int qb_poly_eval_batch(qb_handle* h, int n_polys, const int32_t* poly_ofs, const uint32_t* coef, int n_x, const uint32_t* x, int n_s,
	                       const uint32_t* s, int n_items, const qb_eval_item* items, uint32_t* out)
	{
		return h->e->poly_eval(n_polys, poly_ofs, coef, n_x, x, n_s, s, n_items, items, out);
	}
}
