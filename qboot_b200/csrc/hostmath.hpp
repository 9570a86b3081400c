// qboot_b200 -- host-side helpers on the fixed-limb float: sqrt, ln 2, log, wide division, and the one-time Context
// constants (rho, the rho->z matrix, ln rho).  Host only: these run once per Context, never per table.
//
// Reference for the constants: Context::Context / z_as_func_rho (src/context.cpp:11-26,47-61), power_function and
// RealConverter (src/algebra/real_function.cpp:10-40), lower_triangular_inverse (src/algebra/matrix.cpp:106-123).
// The operation order below is the reference's, so rho and the matrix agree with it bit for bit.
#pragma once

#include <cmath>
#include <vector>

#include "mpf.cuh"
#include "mpf_math.cuh"

namespace qb
{
	// r (LO limbs, `prec` bits) = RN(a / b) for wider operands (N limbs): same long division, rounded once at prec
	template <int LO, int N>
	void div_wide(mpf<LO>& r, const mpf<N>& a, const mpf<N>& b, int prec)
	{
		static_assert(N >= LO, "div_wide");
		// Divide at M = 2N + 2 limbs first, then round that quotient to prec bits.
		constexpr int M = 2 * N + 2;
		mpf<M> aw, bw, qw;
		widen(aw, a);
		widen(bw, b);
		div(qw, aw, bw, 32 * M);
		// qw = RN_{32M}(a/b).  If a/b is exactly representable in 32M bits it is exact; otherwise its distance from any
		// prec-bit rounding boundary exceeds 2^-(32M) relative (|a/b - boundary| >= 1 / (B * 2^(prec+1)) with B < 2^(32N)),
		// so rounding qw to prec bits gives RN_prec(a/b).
		narrow(r, qw, prec);
	}

	// floor(sqrt) based correctly rounded square root
	template <int L>
	void sqrt(mpf<L>& r, const mpf<L>& a, int prec)
	{
		if (is_zero(a))
		{
			copy(r, a);
			return;
		}
		if (kind(a) != K_REG || sign(a))
		{
			if (kind(a) == K_INF && !sign(a))
				copy(r, a);
			else
				set_nan(r);
			return;
		}
		constexpr int N = L + 1;
		// X (2N limbs) = mantissa aligned so that the exponent is even
		uint32_t X[2 * N];
		for (int i = 0; i < 2 * N; ++i) X[i] = 0u;
		int32_t e = a.exp;
		bool odd = (e & 1) != 0;
		// top L limbs <- m (or m >> 1 with the lost bit one limb below)
		for (int i = 0; i < L; ++i) X[2 * N - L + i] = a.m[i];
		if (odd)
		{
			for (int i = 2 * N - L - 1; i < 2 * N - 1; ++i) X[i] = (X[i] >> 1) | (X[i + 1] << 31);
			X[2 * N - 1] >>= 1;
			e += 1;
		}
		// initial approximation by Newton in mpf<N>: y ~ sqrt(0.X)
		mpf<N> x, y, t;
		x.sk = K_REG;
		x.exp = 0;
		for (int i = 0; i < N; ++i) x.m[i] = X[N + i];
		if (odd)
		{
			// renormalise: top bit may be clear
			uint32_t tmpm[N];
			for (int i = 0; i < N; ++i) tmpm[i] = x.m[i];
			round_to<N, N>(x, 0u, 0, tmpm, 0u, 32 * N);
		}
		double xd = to_double(x);
		double yd = std::sqrt(xd);
		// y from double
		{
			int ex;
			double fr = std::frexp(yd, &ex);
			uint64_t mant = uint64_t(std::ldexp(fr, 63));
			uint32_t Y[2] = {uint32_t(mant), uint32_t(mant >> 32)};
			round_to<N, 2>(y, 0u, ex + 1, Y, 0u, 32 * N);
		}
		for (int it = 0; it < 8; ++it)
		{
			div(t, x, y, 32 * N);
			add(y, y, t, 32 * N);
			y.exp -= 1;
		}
		// integer candidate R = floor(y * 2^(32N)) (y in [1/2, 1))
		uint32_t R[N];
		{
			// y = 0.m * 2^exp with exp in {0, ...}; shift mantissa right by -exp (exp <= 0 expected, = 0 typically)
			int sh = -y.exp;
			for (int i = 0; i < N; ++i) R[i] = y.m[i];
			while (sh > 0)
			{
				for (int i = 0; i < N - 1; ++i) R[i] = (R[i] >> 1) | (R[i + 1] << 31);
				R[N - 1] >>= 1;
				--sh;
			}
			if (y.exp > 0)
			{
				// y == 1.0 (can only arise from rounding): use all ones
				for (int i = 0; i < N; ++i) R[i] = 0xFFFFFFFFu;
			}
		}
		// fix up: largest R with R^2 <= X
		auto square = [](const uint32_t* Rv, uint32_t* out) { mul_limbs<N, N>(out, Rv, Rv); };
		auto cmp2 = [](const uint32_t* A, const uint32_t* B) {
			for (int i = 2 * N - 1; i >= 0; --i)
				if (A[i] != B[i]) return A[i] < B[i] ? -1 : 1;
			return 0;
		};
		uint32_t SQ[2 * N];
		for (int guard = 0; guard < 64; ++guard)
		{
			square(R, SQ);
			if (cmp2(SQ, X) > 0)
			{
				// R -= 1
				for (int i = 0; i < N; ++i)
					if (R[i]-- != 0u) break;
				continue;
			}
			// check (R+1)^2 <= X  <=>  SQ + 2R + 1 <= X
			uint32_t T[2 * N];
			uint64_t c = 1;
			for (int i = 0; i < 2 * N; ++i)
			{
				uint64_t add2 = 0;
				if (i < N) add2 = (uint64_t(R[i]) << 1) & 0xFFFFFFFFull;
				if (i >= 1 && i <= N) add2 |= (R[i - 1] >> 31);
				c += uint64_t(SQ[i]) + add2;
				T[i] = uint32_t(c);
				c >>= 32;
			}
			bool all_ones = true;
			for (int i = 0; i < N; ++i) all_ones &= (R[i] == 0xFFFFFFFFu);
			if (!c && !all_ones && cmp2(T, X) <= 0)
			{
				for (int i = 0; i < N; ++i)
					if (++R[i] != 0u) break;
				continue;
			}
			break;
		}
		square(R, SQ);
		uint32_t st = cmp2(SQ, X) != 0 ? 1u : 0u;
		round_to<L, N>(r, 0u, e / 2, R, st, prec);
	}

	// ln 2 to 32E bits:  ln 2 = 18 atanh(1/26) - 2 atanh(1/4801) + 8 atanh(1/8749)
	template <int E>
	void const_ln2(mpf<E>& r)
	{
		constexpr int W = E + 1;  // one guard limb
		const int pw = 32 * W;
		auto atanh_inv = [&](mpf<W>& out, int64_t n) {
			mpf<W> term, pwr, t;
			set_si(pwr, 1, pw);
			div_si(pwr, pwr, n, pw);  // 1/n
			copy(out, pwr);
			int64_t n2 = n * n;
			for (int k = 1; k < 40 * W; ++k)
			{
				div_si(pwr, pwr, n2, pw);
				if (is_zero(pwr) || pwr.exp < -pw - 8) break;
				div_si(t, pwr, 2 * k + 1, pw);
				add(out, out, t, pw);
			}
			(void)term;
		};
		mpf<W> a, b, c, s;
		atanh_inv(a, 26);
		atanh_inv(b, 4801);
		atanh_inv(c, 8749);
		mul_si(a, a, 18, pw);
		mul_si(b, b, 2, pw);
		mul_si(c, c, 8, pw);
		sub(s, a, b, pw);
		add(s, s, c, pw);
		narrow(r, s, 32 * E);
	}

	// ln x to ~32E bits for x > 0 given at any width <= E (Newton on exp: y += 2 (x - e^y) / (x + e^y))
	template <int E>
	void log_ext(mpf<E>& r, const mpf<E>& x, const mpf<E>& ln2)
	{
		constexpr int W = E + 2;
		const int pw = 32 * W;
		mpf<W> xw, l2w, y, ey, num, den, t;
		widen(xw, x);
		// ln2 at W limbs
		const_ln2(l2w);
		(void)ln2;
		double xd = to_double(x);
		// guard against double overflow for large exponents: ln x = ln(m) + exp * ln 2
		double md;
		{
			mpf<E> xm;
			copy(xm, x);
			xm.exp = 0;
			md = to_double(xm);
		}
		(void)xd;
		double yd = std::log(md) + double(x.exp) * 0.6931471805599453;
		{
			int ex;
			double fr = std::frexp(std::fabs(yd), &ex);
			uint64_t mant = uint64_t(std::ldexp(fr, 63));
			uint32_t Y[2] = {uint32_t(mant), uint32_t(mant >> 32)};
			if (mant == 0)
				set_zero(y);
			else
				round_to<W, 2>(y, yd < 0 ? SIGN_BIT : 0u, ex + 1, Y, 0u, pw);
		}
		for (int it = 0; it < 7; ++it)
		{
			exp_ext(ey, y, l2w);
			sub(num, xw, ey, pw);
			add(den, xw, ey, pw);
			div(t, num, den, pw);
			t.exp += (kind(t) == K_REG) ? 1 : 0;
			add(y, y, t, pw);
		}
		narrow(r, y, 32 * E);
	}

	// One-time constants of a Context
	template <int L>
	struct ContextConsts
	{
		static constexpr int E = L + QB_EXT_LIMBS;
		mpf<L> rho;
		std::vector<mpf<L>> mat;    // (lambda+1)^2 row-major: rho-Taylor -> z-Taylor
		std::vector<mpf<L>> kfact;  // k!
		mpf<E> ln2e, ln4e, lnrhoe;
		std::vector<mpf<E>> rhoinv;  // rho^-k at the extended width
		// the logarithms once more at the second Ziv level (mpf_math.cuh)
		static constexpr int E2 = L + QB_EXT2_LIMBS;
		mpf<E2> ln2w, ln4w, lnrhow;
	};

	namespace detail
	{
		// truncated product of two series (RealFunction mul, real_function.hpp:121-128)
		template <int L>
		void series_mul(std::vector<mpf<L>>& z, const std::vector<mpf<L>>& x, const std::vector<mpf<L>>& y, int prec)
		{
			int lam = int(x.size()) - 1;
			z.assign(size_t(lam) + 1, mpf<L>{});
			for (auto& v : z) set_zero(v);
			mpf<L> t;
			for (int k1 = 0; k1 <= lam; ++k1)
				for (int k2 = 0; k1 + k2 <= lam; ++k2)
				{
					mul(t, x[size_t(k1)], y[size_t(k2)], prec);
					add(z[size_t(k1 + k2)], z[size_t(k1 + k2)], t, prec);
				}
		}
		// RealConverter(func) (real_function.cpp:20-31): row n = coefficients of func^n
		template <int L>
		void converter(std::vector<mpf<L>>& mat, const std::vector<mpf<L>>& func, int prec)
		{
			int lam = int(func.size()) - 1, n1 = lam + 1;
			mat.assign(size_t(n1) * n1, mpf<L>{});
			for (auto& v : mat) set_zero(v);
			set_si(mat[0], 1, prec);
			std::vector<mpf<L>> pf = func, nx;
			for (int n = 1; n <= lam; ++n)
			{
				for (int i = n; i <= lam; ++i) copy(mat[size_t(n) * n1 + i], pf[size_t(i)]);
				series_mul(nx, pf, func, prec);
				pf.swap(nx);
			}
		}
	}  // namespace detail

	template <int L>
	void make_context_consts(ContextConsts<L>& c, uint32_t lambda, int prec)
	{
		constexpr int E = ContextConsts<L>::E;
		const int lam = int(lambda), n1 = lam + 1;
		mpf<L> eight, sqrt8, a, tmp, t, u;
		set_si(eight, 8, prec);
		sqrt(sqrt8, eight, prec);
		si_sub(c.rho, 3, sqrt8, prec);  // rho = 3 - sqrt(8)
		// f = (a + r)^-2, a = 4 - sqrt(8)                                         (power_function)
		si_sub(a, 4, sqrt8, prec);
		std::vector<mpf<L>> f(static_cast<size_t>(n1));
		{
			// f[0] = pow(a, -2) correctly rounded = RN(1 / a^2): a^2 exact in 2L limbs
			mpf<2 * L> a2, one2;
			uint32_t am[L], P2[2 * L];
			for (int i = 0; i < L; ++i) am[i] = a.m[i];
			mul_limbs<L, L>(P2, am, am);
			round_to<2 * L, 2 * L>(a2, 0u, 2 * a.exp, P2, 0u, 64 * L);  // exact
			set_si(one2, 1, 64 * L);
			div_wide<L, 2 * L>(f[0], one2, a2, prec);
		}
		mpf<L> one;
		set_si(one, 1, prec);
		div(tmp, one, a, prec);  // b / a with b = 1
		for (int k = 1; k <= lam; ++k)
		{
			mul(t, f[size_t(k - 1)], tmp, prec);
			set_si(u, -2, prec);
			add_si(u, u, -int64_t(k - 1), prec);  // p - (k - 1)
			mul(t, t, u, prec);
			div_si(f[size_t(k)], t, int64_t(k), prec);
		}
		// z - 1/2 = sqrt(8) r f - r^2 f / 2                                       (z_as_func_rho)
		const size_t nn = size_t(n1);
		std::vector<mpf<L>> g1(nn), g2(nn), func(nn);
		for (int i = 0; i <= lam; ++i)
		{
			if (i >= 1)
				mul(g1[size_t(i)], f[size_t(i - 1)], sqrt8, prec);
			else
			{
				set_zero(g1[0]);
				mul(g1[0], g1[0], sqrt8, prec);
			}
			if (i >= 2)
				div_si(g2[size_t(i)], f[size_t(i - 2)], -2, prec);
			else
			{
				set_zero(g2[size_t(i)]);
				div_si(g2[size_t(i)], g2[size_t(i)], -2, prec);
			}
			add(func[size_t(i)], g1[size_t(i)], g2[size_t(i)], prec);
		}
		// RealConverter(func).inverse()
		std::vector<mpf<L>> m0, inv(size_t(n1) * n1), res(size_t(n1) * n1);
		detail::converter(m0, func, prec);
		for (int i = 0; i < n1; ++i)
			for (int j = 0; j < n1; ++j) copy(inv[size_t(i) * n1 + j], m0[size_t(j) * n1 + i]);  // transpose
		for (auto& v : res) set_zero(v);
		mpf<L> s;
		for (int i = 0; i < n1; ++i)
		{
			div(res[size_t(i) * n1 + i], one, inv[size_t(i) * n1 + i], prec);
			for (int j = 0; j < i; ++j)
			{
				set_zero(s);
				for (int k = j; k < i; ++k)
				{
					mul(t, inv[size_t(i) * n1 + k], res[size_t(k) * n1 + j], prec);
					add(s, s, t, prec);
				}
				neg(s);
				div(res[size_t(i) * n1 + j], s, inv[size_t(i) * n1 + i], prec);
			}
		}
		std::vector<mpf<L>> fin(nn);
		set_zero(fin[0]);
		for (int k = 1; k <= lam; ++k) copy(fin[size_t(k)], res[size_t(k) * n1 + 1]);
		detail::converter(c.mat, fin, prec);
		// k!
		c.kfact.resize(size_t(n1));
		set_si(c.kfact[0], 1, prec);
		for (int k = 1; k <= lam; ++k) mul_si(c.kfact[size_t(k)], c.kfact[size_t(k - 1)], int64_t(k), prec);
		// logs in the extended format
		const_ln2(c.ln2e);
		copy(c.ln4e, c.ln2e);
		c.ln4e.exp += 1;
		mpf<E> rw;
		widen(rw, c.rho);
		log_ext(c.lnrhoe, rw, c.ln2e);
		{
			constexpr int E2 = ContextConsts<L>::E2;
			const_ln2(c.ln2w);
			copy(c.ln4w, c.ln2w);
			c.ln4w.exp += 1;
			mpf<E2> rw2;
			widen(rw2, c.rho);
			log_ext(c.lnrhow, rw2, c.ln2w);
		}
		// rho^-k = exp(-k ln rho), each from its own exponential (no error accumulation)
		c.rhoinv.resize(size_t(n1));
		for (int k = 0; k <= lam; ++k)
		{
			mpf<E> tk;
			mul_si(tk, c.lnrhoe, -int64_t(k), 32 * E);
			exp_ext(c.rhoinv[size_t(k)], tk, c.ln2e);
		}
	}
}  // namespace qb
