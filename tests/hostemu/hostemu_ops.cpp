// TEST HARNESS ONLY: compiles the engine's `__host__ __device__` multiprecision routines (qboot_b200/csrc/mpf.cuh)
// for the host so that the CPU test-suite (-m "not gpu") can pin them bit-for-bit against MPFR (oracle/_ref).
// Nothing in the product links or loads this file.
#include <cstdint>
#include <cstring>

#include "../../qboot_b200/csrc/polyfx.cuh"
#include "../../qboot_b200/csrc/mpf.cuh"
#include "../../qboot_b200/csrc/mpf_math.cuh"
#include "../../qboot_b200/csrc/fastops.cuh"

namespace
{
	template <int L>
	int run(int prec, int op, int n, const uint32_t* a, const uint32_t* b, const uint32_t* c, const int64_t* ia,
	        const int64_t* ib, uint32_t* out, const uint32_t* consts)
	{
		using T = qb::mpf<L>;
		static_assert(sizeof(T) == 4 * (L + 2), "layout");
		const T* A = reinterpret_cast<const T*>(a);
		const T* B = reinterpret_cast<const T*>(b);
		const T* Cc = reinterpret_cast<const T*>(c);
		T* O = reinterpret_cast<T*>(out);
		T zero;
		qb::set_zero(zero);
		for (int i = 0; i < n; ++i)
		{
			const T& x = a ? A[i] : zero;
			const T& y = b ? B[i] : zero;
			const T& z = c ? Cc[i] : zero;
			int64_t p = ia ? ia[i] : 0, q = ib ? ib[i] : 1;
			T r;
			qb::set_zero(r);
			switch (op)
			{
			case 0: qb::add(r, x, y, prec); break;
			case 1: qb::sub(r, x, y, prec); break;
			case 2: qb::mul(r, x, y, prec); break;
			case 3: qb::div(r, x, y, prec); break;
			case 4: qb::fma(r, x, y, z, prec); break;
			case 6: qb::mul_si(r, x, p, prec); break;
			case 7: qb::add_si(r, x, p, prec); break;
			case 8: qb::div_si(r, x, p, prec); break;
			case 9: qb::mul_q(r, x, p, uint32_t(q), prec); break;
			case 10: qb::div_q(r, x, p, uint32_t(q), prec); break;
			case 11: qb::add_q(r, x, p, uint32_t(q), prec); break;
			case 12: qb::set_q(r, p, uint32_t(q), prec); break;
			case 18: qb::si_sub(r, p, x, prec); break;
			case 20: qb::sub_q(r, x, p, uint32_t(q), prec); break;
			case 22: qb::set_ui_2exp(r, uint64_t(p), int32_t(q), prec); break;
			case 23: r.sk = uint32_t(qb::cmp(x, y) + 1); break;
			case 24: r.sk = qb::is_integer(x) ? 1u : 0u; break;
			case 40: qb::fast::mul(r, x, y, prec); break;
			case 41: qb::fast::fma(r, x, y, z, prec); break;
			case 42: qb::fast::add(r, x, y, prec); break;
			case 43: qb::fast::add(r, x, y, prec, true); break;
			case 44: qb::fast::mul_small(r, x, p, prec); break;
			case 45: qb::fast::add_small(r, x, p, prec); break;
			case 46: qb::fast::div(r, x, y, prec); break;
			case 47:  // one Horner step with an integer argument: RN(RN(x * p) + y)
				qb::copy(r, x);
				if ((x.sk & 3u) != qb::K_REG || (y.sk & 3u) != qb::K_REG || p <= 0 || (p >> 32) != 0 ||
				    !qb::fast::horner_step(r, uint32_t(p), y, prec))
				{
					qb::fast::mul_small(r, x, p, prec);
					qb::fast::add(r, r, y, prec);
					r.sk |= 0x100u;  // marks "fast path declined" for the test's statistics (cleared by the harness)
				}
				break;
			case 13:  // pow(x, y) with x > 0 given ln(x) in consts (extended precision)
			case 14:  // ui_pow
			case 15:  // exp
				// false: no constants supplied, or (13 / 14) the rounding is in doubt at this width -- r is set all the same
				if (!qb::hostemu_math<L>(prec, op, x, y, p, consts, r))
				{
					O[i] = r;
					return -2;
				}
				break;
			default: return -1;
			}
			O[i] = r;
		}
		return 0;
	}
}  // namespace

namespace
{
	// the fixed-point evaluation of one polynomial (csrc/polyfx.cuh) at n_count integers; status 1 = fixed point, 0 = the
	// floating-point Horner fall-back (frame not exact for these coefficients, or a value left the window)
	template <int L>
	int poly_fx(int prec, int deg, int nbits, const uint32_t* coefs, int n_count, const uint32_t* ns, uint32_t* out, int32_t* status)
	{
		using T = qb::mpf<L>;
		constexpr int W = L + 2;
		const T* c = reinterpret_cast<const T*>(coefs);
		T* O = reinterpret_cast<T*>(out);
		int32_t e[8];
		bool have[8];
		for (int d = 0; d <= deg; ++d)
		{
			have[d] = qb::kind(c[d]) == qb::K_REG;
			e[d] = c[d].exp;
		}
		const int32_t E = qb::fx::frame_exponent(e, have, deg, nbits);
		uint32_t C[8][W], cs[8];
		bool ok = true;
		for (int d = 0; d <= deg; ++d) ok = qb::fx::from_record<L>(C[d], cs[d], c[d], E) && ok;
		for (int i = 0; i < n_count; ++i)
		{
			bool good = ok;
			T p;
			if (good)
			{
				uint32_t V[W], sg = cs[deg];
				for (int j = 0; j < W; ++j) V[j] = C[deg][j];
				for (int d = deg - 1; d >= 0 && good; --d) good = qb::fx::horner_step<L>(V, sg, ns[i], C[d], cs[d], prec);
				if (good) good = qb::fx::to_record<L>(p, V, sg, E);
			}
			if (!good)
			{
				qb::copy(p, c[deg]);
				for (int d = deg - 1; d >= 0; --d)
				{
					qb::mul_si(p, p, int64_t(ns[i]), prec);
					qb::add(p, p, c[d], prec);
				}
			}
			O[i] = p;
			status[i] = good ? 1 : 0;
		}
		return 0;
	}
}  // namespace

#define QB_FOR_L(X) X(2) X(3) X(4) X(5) X(8) X(16) X(19) X(25) X(32) X(38) X(48)

extern "C" int qbh_op(int prec, int op, int n, const uint32_t* a, const uint32_t* b, const uint32_t* c,
                      const int64_t* ia, const int64_t* ib, uint32_t* out, const uint32_t* consts)
{
	int L = (prec + 31) / 32;
	switch (L)
	{
#define X(LL) \
	case LL: return run<LL>(prec, op, n, a, b, c, ia, ib, out, consts);
		QB_FOR_L(X)
#undef X
	default: return -100;
	}
}

extern "C" int qbh_poly_fx(int prec, int deg, int nbits, const uint32_t* coefs, int n_count, const uint32_t* ns, uint32_t* out, int32_t* status)
{
	int L = (prec + 31) / 32;
	if (deg < 0 || deg > 7) return -1;
	switch (L)
	{
#define X(LL) \
	case LL: return poly_fx<LL>(prec, deg, nbits, coefs, n_count, ns, out, status);
		QB_FOR_L(X)
#undef X
	default: return -100;
	}
}
