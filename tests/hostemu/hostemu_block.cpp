// TEST HARNESS ONLY: host build of the engine's per-table `__host__ __device__` stage functions, so the CPU suite can
// pin each stage against the reference (oracle/_ref) without a GPU.  Nothing in the product links or loads this file.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../qboot_b200/csrc/block.cuh"
#include "../../qboot_b200/csrc/hor_table.inc"
#include "../../qboot_b200/csrc/hostmath.hpp"

#define QB_FOR_L(X) X(2) X(3) X(4) X(5) X(8) X(16) X(19) X(25) X(32) X(38) X(48)

namespace
{
	template <int L>
	int rec_coeffs_t(int prec, int64_t en, uint32_t ed, const uint32_t* d, uint32_t spin, const uint32_t* S,
	                 const uint32_t* P, uint32_t* out)
	{
		using T = qb::mpf<L>;
		const T& D = *reinterpret_cast<const T*>(d);
		const T& s = *reinterpret_cast<const T*>(S);
		const T& p = *reinterpret_cast<const T*>(P);
		qb::rec_coeffs<L>(reinterpret_cast<T*>(out), D, spin, s, p, qb::Rational{en, ed}, prec,
		                  spin == 0 ? QB_HOR_POLY_0 : QB_HOR_POLY_1, spin == 0 ? QB_HOR_MONO_0 : QB_HOR_MONO_1);
		return spin == 0 ? 6 : 8;
	}
	template <int L>
	int hblock_t(int prec, uint32_t n_max, int64_t en, uint32_t ed, const uint32_t* d, uint32_t spin, const uint32_t* S,
	             const uint32_t* P, const uint32_t* b0, uint32_t* out)
	{
		using T = qb::mpf<L>;
		T coef[40];
		rec_coeffs_t<L>(prec, en, ed, d, spin, S, P, reinterpret_cast<uint32_t*>(coef));
		T* b = reinterpret_cast<T*>(out);
		b[0] = *reinterpret_cast<const T*>(b0);
		qb::series<L>(b, 1, n_max, coef, spin, prec);
		return 0;
	}
	// full table, staged exactly like the device pipeline (series -> scale by 4^Delta -> lambda+1 Horner chains ->
	// rho->z -> transverse recursion).  consts_ext = [ln2, ln4, lnrho] in E = L + QB_EXT_LIMBS limbs.
	template <int L>
	int gblock_t(int prec, uint32_t n_max, uint32_t lambda, int64_t en, uint32_t ed, const uint32_t* d, uint32_t spin,
	             const uint32_t* S, const uint32_t* P, const uint32_t* b0, const uint32_t* rho_w, const uint32_t* mat,
	             const uint32_t* consts_ext, uint32_t* out, uint32_t* out_real)
	{
		using T = qb::mpf<L>;
		constexpr int E = L + QB_EXT_LIMBS;
		using TE = qb::mpf<E>;
		const TE* ce = reinterpret_cast<const TE*>(consts_ext);
		const T& D = *reinterpret_cast<const T*>(d);
		const T& s = *reinterpret_cast<const T*>(S);
		const T& p = *reinterpret_cast<const T*>(P);
		const T& rho = *reinterpret_cast<const T*>(rho_w);
		const T* M = reinterpret_cast<const T*>(mat);
		qb::Rational eps{en, ed};
		std::vector<T> b(n_max + 1);
		hblock_t<L>(prec, n_max, en, ed, d, spin, S, P, b0, reinterpret_cast<uint32_t*>(b.data()));
		T c4;
		qb::pow_from_log<L>(c4, ce[1], D, ce[0], prec);
		for (uint32_t n = 0; n <= n_max; ++n) qb::mul(b[n], b[n], c4, prec);
		std::vector<T> fr(lambda + 1);
		T q, kf, sacc, rp;
		qb::copy(q, D);
		qb::set_si(kf, 1, prec);
		for (uint32_t k = 0; k <= lambda; ++k)
		{
			if (k >= 1)
			{
				qb::mul_si(kf, kf, int64_t(k), prec);
				T qprev;
				qb::copy(qprev, q);
				for (uint32_t n = 0; n <= n_max; ++n) qb::deriv_scale(b[n], b[n], qprev, n, prec);
				qb::add_si(q, q, -1, prec);
			}
			qb::set_zero(sacc);
			for (uint32_t n = n_max + 1; n-- > 0;) qb::deriv_horner(sacc, rho, b[n], prec);
			qb::pow_from_log<L>(rp, ce[2], q, ce[0], prec);
			qb::deriv_finish(fr[k], sacc, rp, kf, k, prec);
		}
		if (out_real) std::memcpy(out_real, fr.data(), sizeof(T) * (lambda + 1));
		T* f = reinterpret_cast<T*>(out);
		for (int i = 0; i < qb::cf_dim(int(lambda)); ++i) qb::set_zero(f[i]);
		for (uint32_t i = 0; i <= lambda; ++i) qb::rho_to_z_entry(f[i], fr.data(), 1, M, int(lambda), int(i), prec);
		qb::expand_off_diag(f, 1, int(lambda), D, spin, s, p, eps, prec);
		return qb::cf_dim(int(lambda));
	}
	template <int L>
	int fblock_t(int prec, uint32_t lambda, const uint32_t* dw, const uint32_t* g, int sym, const uint32_t* consts_ext,
	             uint32_t* out)
	{
		using T = qb::mpf<L>;
		constexpr int E = L + QB_EXT_LIMBS;
		using TE = qb::mpf<E>;
		const TE* ce = reinterpret_cast<const TE*>(consts_ext);
		const T& d = *reinterpret_cast<const T*>(dw);
		T md, p4;
		qb::copy(md, d);
		qb::neg(md);
		qb::pow_from_log<L>(p4, ce[1], md, ce[0], prec);
		std::vector<T> v(qb::cf_dim(int(lambda)));
		qb::v_to_d_table(v.data(), 1, int(lambda), d, p4, prec);
		T* o = reinterpret_cast<T*>(out);
		int cnt = 0;
		for (int N = 0; N <= int(lambda) / 2; ++N)
			for (int Mx = 0; Mx + 2 * N <= int(lambda); ++Mx)
				if ((Mx & 1) == sym)
				{
					qb::fblock_entry(o[qb::cf_index_sym(int(lambda), sym, Mx, N)], v.data(), 1,
					                 reinterpret_cast<const T*>(g), 1, int(lambda), Mx, N, prec);
					++cnt;
				}
		return cnt;
	}
	template <int L>
	int consts_t(int prec, uint32_t lambda, uint32_t* rho, uint32_t* mat, uint32_t* kfact, uint32_t* ext3)
	{
		qb::ContextConsts<L> c;
		qb::make_context_consts<L>(c, lambda, prec);
		std::memcpy(rho, &c.rho, sizeof(c.rho));
		std::memcpy(mat, c.mat.data(), sizeof(qb::mpf<L>) * c.mat.size());
		std::memcpy(kfact, c.kfact.data(), sizeof(qb::mpf<L>) * c.kfact.size());
		constexpr int E = L + QB_EXT_LIMBS;
		std::memcpy(ext3, &c.ln2e, sizeof(qb::mpf<E>));
		std::memcpy(ext3 + (E + 2), &c.ln4e, sizeof(qb::mpf<E>));
		std::memcpy(ext3 + 2 * (E + 2), &c.lnrhoe, sizeof(qb::mpf<E>));
		return 0;
	}
	template <int L>
	int sqrt_t(int prec, int n, const uint32_t* a, uint32_t* out)
	{
		using T = qb::mpf<L>;
		for (int i = 0; i < n; ++i) qb::sqrt(reinterpret_cast<T*>(out)[i], reinterpret_cast<const T*>(a)[i], prec);
		return 0;
	}
}  // namespace

extern "C"
{
	int qbh_context_consts(int prec, uint32_t lambda, uint32_t* rho, uint32_t* mat, uint32_t* kfact, uint32_t* ext3)
	{
		switch ((prec + 31) / 32)
		{
#define X(LL) \
	case LL: return consts_t<LL>(prec, lambda, rho, mat, kfact, ext3);
			QB_FOR_L(X)
#undef X
		default: return -100;
		}
	}
	int qbh_sqrt(int prec, int n, const uint32_t* a, uint32_t* out)
	{
		switch ((prec + 31) / 32)
		{
#define X(LL) \
	case LL: return sqrt_t<LL>(prec, n, a, out);
			QB_FOR_L(X)
#undef X
		default: return -100;
		}
	}
	int qbh_gblock(int prec, uint32_t n_max, uint32_t lambda, int64_t en, uint32_t ed, const uint32_t* d, uint32_t spin,
	               const uint32_t* S, const uint32_t* P, const uint32_t* b0, const uint32_t* rho, const uint32_t* mat,
	               const uint32_t* consts_ext, uint32_t* out, uint32_t* out_real)
	{
		switch ((prec + 31) / 32)
		{
#define X(LL) \
	case LL: return gblock_t<LL>(prec, n_max, lambda, en, ed, d, spin, S, P, b0, rho, mat, consts_ext, out, out_real);
			QB_FOR_L(X)
#undef X
		default: return -100;
		}
	}
	int qbh_fblock(int prec, uint32_t lambda, const uint32_t* d, const uint32_t* g, int sym, const uint32_t* consts_ext,
	               uint32_t* out)
	{
		switch ((prec + 31) / 32)
		{
#define X(LL) \
	case LL: return fblock_t<LL>(prec, lambda, d, g, sym, consts_ext, out);
			QB_FOR_L(X)
#undef X
		default: return -100;
		}
	}
	int qbh_rec_coeffs(int prec, int64_t en, uint32_t ed, const uint32_t* d, uint32_t spin, const uint32_t* S,
	                   const uint32_t* P, uint32_t* out)
	{
		switch ((prec + 31) / 32)
		{
#define X(LL) \
	case LL: return rec_coeffs_t<LL>(prec, en, ed, d, spin, S, P, out);
			QB_FOR_L(X)
#undef X
		default: return -100;
		}
	}
	int qbh_hblock(int prec, uint32_t n_max, int64_t en, uint32_t ed, const uint32_t* d, uint32_t spin,
	               const uint32_t* S, const uint32_t* P, const uint32_t* b0, uint32_t* out)
	{
		switch ((prec + 31) / 32)
		{
#define X(LL) \
	case LL: return hblock_t<LL>(prec, n_max, en, ed, d, spin, S, P, b0, out);
			QB_FOR_L(X)
#undef X
		default: return -100;
		}
	}
}
