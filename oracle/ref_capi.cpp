// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or called from the product path.
//
// C-ABI driver over the UNMODIFIED reference (qboot v0.8.1, compiled from /root/reference by oracle/Makefile into
// oracle/_ref/libqboot_ref.so).  tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
// load it with ctypes to obtain ground-truth values and CPU timings.  Every entry point calls the reference's public
// C++ API (cited per function) and only converts between mpfr_t and the flat "qb words" interchange format:
//
//   one number = W = L + 2 uint32 words, L = ceil(prec / 32)
//     w[0]  kind | sign << 31     kind: 0 zero, 1 regular, 2 inf, 3 nan
//     w[1]  exponent (int32), value = 0.1xxxx (binary) * 2^exp  (MPFR convention), 0 for zero
//     w[2 .. 2 + L)  mantissa limbs, least significant first, left aligned (bit 31 of w[L + 1] is set)
//
#include <chrono>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <optional>
#include <sstream>
#include <string>
#include <tuple>
#include <vector>

#include "qboot/qboot.hpp"

using qboot::mp::rational;
using qboot::mp::real;
namespace alg = qboot::algebra;

namespace
{
	struct real_peek
	{
		// `real` holds exactly one mpfr_t (include/qboot/mp/real.hpp:87-90); peek through a layout-compatible view.
		static mpfr_ptr get(real& r) { return reinterpret_cast<mpfr_ptr>(&r); }
		static mpfr_srcptr get(const real& r) { return reinterpret_cast<mpfr_srcptr>(&r); }
	};
	static_assert(sizeof(real) == sizeof(__mpfr_struct), "real must wrap a single mpfr_t");

	inline uint32_t nlimbs(int prec) { return uint32_t((prec + 31) / 32); }

	void to_words(mpfr_srcptr x, int prec, uint32_t* w)
	{
		uint32_t L = nlimbs(prec);
		std::memset(w, 0, sizeof(uint32_t) * (L + 2));
		uint32_t sign = mpfr_signbit(x) ? 0x80000000u : 0u;
		if (mpfr_nan_p(x))
		{
			w[0] = 3u;
			return;
		}
		if (mpfr_inf_p(x))
		{
			w[0] = 2u | sign;
			return;
		}
		if (mpfr_zero_p(x))
		{
			w[0] = sign;
			return;
		}
		w[0] = 1u | sign;
		w[1] = uint32_t(int32_t(x->_mpfr_exp));
		// MPFR stores ceil(p / 64) 64-bit limbs, mantissa left aligned in the top limb
		long p = x->_mpfr_prec;
		long n64 = (p + 63) / 64;
		// view as 32-bit words little endian: word j of the 2*n64 words
		// our L words are the TOP L words of that array
		for (uint32_t i = 0; i < L; ++i)
		{
			long j = 2 * n64 - long(L) + long(i);  // index into 32-bit view
			if (j < 0) continue;
			uint64_t limb = x->_mpfr_d[j / 2];
			w[2 + i] = uint32_t((j & 1) ? (limb >> 32) : limb);
		}
	}
	void from_words(const uint32_t* w, int prec, mpfr_ptr x)
	{
		uint32_t L = nlimbs(prec);
		uint32_t kind = w[0] & 3u;
		int sign = (w[0] >> 31) ? -1 : 1;
		if (kind == 3u)
		{
			mpfr_set_nan(x);
			return;
		}
		if (kind == 2u)
		{
			mpfr_set_inf(x, sign);
			return;
		}
		if (kind == 0u)
		{
			mpfr_set_zero(x, sign);
			return;
		}
		// build through an mpz to stay independent of the limb layout
		mpz_t z;
		mpz_init(z);
		for (uint32_t i = L; i-- > 0;)
		{
			mpz_mul_2exp(z, z, 32);
			mpz_add_ui(z, z, w[2 + i]);
		}
		if (sign < 0) mpz_neg(z, z);
		mpfr_set_z_2exp(x, z, long(int32_t(w[1])) - 32L * long(L), MPFR_RNDN);
		mpz_clear(z);
	}
	real mk(const uint32_t* w, int prec)
	{
		real r;
		from_words(w, prec, real_peek::get(r));
		return r;
	}
	void put(const real& r, int prec, uint32_t* w) { to_words(real_peek::get(r), prec, w); }
	void set_prec(int prec)
	{
		qboot::mp::global_prec = prec;
		qboot::mp::global_rnd = MPFR_RNDN;
	}

	std::mutex ctx_mutex;
	std::map<std::tuple<int, uint32_t, uint32_t, std::string, uint32_t>, std::unique_ptr<qboot::Context>> ctx_cache;
	const qboot::Context& get_ctx(int prec, uint32_t n_max, uint32_t lambda, const char* dim, uint32_t p = 1)
	{
		std::lock_guard<std::mutex> g(ctx_mutex);
		auto key = std::make_tuple(prec, n_max, lambda, std::string(dim), p);
		auto it = ctx_cache.find(key);
		if (it == ctx_cache.end())
		{
			set_prec(prec);
			it = ctx_cache.emplace(key, std::make_unique<qboot::Context>(n_max, lambda, rational(dim), p)).first;
		}
		return *it->second;
	}
	rational eps_of(const char* dim) { return rational(dim) / 2 - 1; }

	using dict = std::map<std::string, rational, std::less<>>;
	using Op = qboot::PrimaryOperator;
	constexpr auto Odd = alg::FunctionSymmetry::Odd;
	constexpr auto Even = alg::FunctionSymmetry::Even;

	// Workload definitions (the reference ships them only as driver programs, sample/main.cpp:32-59 and
	// test/src/main.cpp:97-208; restated here against the same public API so they can be called as a library).
	// `finite` adds the finite-range operators [lb, ub) of the drivers.
	qboot::BootstrapEquation model_single(const qboot::Context& c, const dict& d, uint32_t numax, uint32_t maxspin,
	                                      bool finite)
	{
		Op sig(real(d.at("s")), 0, c), eps(real(d.at("e")), 0, c), unit(c), stress(2, c);
		std::array ssss{sig, sig, sig, sig};
		std::vector<qboot::Sector> secs{{"unit", 1, {real(1)}}, {"e", 1}, {"T", 1}};
		qboot::Sector even("even", 1, qboot::SectorType::Continuous);
		if (finite) even.add_op(0, real("1.409"), real("1.4135"));
		even.add_op(0, real(6));
		even.add_op(2, real(5));
		for (uint32_t l = 4; l <= maxspin; l += 2) even.add_op(l);
		secs.push_back(std::move(even));
		qboot::BootstrapEquation boot(c, std::move(secs), numax);
		qboot::Equation eq(boot, Odd);
		eq.add("unit", unit, ssss);
		eq.add("e", eps, ssss);
		eq.add("T", stress, ssss);
		eq.add("even", ssss);
		boot.add_equation(std::move(eq));
		boot.finish();
		return boot;
	}
	qboot::BootstrapEquation model_mixed(const qboot::Context& c, const dict& d, uint32_t numax, uint32_t maxspin,
	                                     bool finite)
	{
		Op s(real(d.at("s")), 0, c), e(real(d.at("e")), 0, c), e1(real(d.at("e1")), 0, c), unit(c), stress(2, c);
		std::array ssss{s, s, s, s}, eeee{e, e, e, e}, eses{e, s, e, s}, eess{e, e, s, s}, esse{e, s, s, e};
		std::vector<qboot::Sector> secs{{"unit", 1, {real(1)}}, {"scalar", 2}, {"e1", 2}, {"T", 2, {s.delta(), e.delta()}}};
		qboot::Sector even("even", 2, qboot::SectorType::Continuous);
		if (finite) even.add_op(0, real("3.815"), real("3.845"));
		even.add_op(0, real(6));
		even.add_op(2, real(5));
		for (uint32_t l = 4; l <= maxspin; l += 2) even.add_op(l);
		secs.push_back(even);
		qboot::Sector oddp("odd+", 1, qboot::SectorType::Continuous);
		oddp.add_op(0, real(3));
		if (finite) oddp.add_op(0, real("5.25"), real("5.33"));
		for (uint32_t l = 2; l <= maxspin; l += 2) oddp.add_op(l);
		secs.push_back(oddp);
		qboot::Sector oddm("odd-", 1, qboot::SectorType::Continuous);
		for (uint32_t l = 1; l <= maxspin; l += 2) oddm.add_op(l);
		secs.push_back(oddm);
		qboot::BootstrapEquation boot(c, secs, numax);
		{
			qboot::Equation eq(boot, Odd);
			eq.add("unit", unit, ssss);
			eq.add("scalar", 0, 0, e, ssss);
			eq.add("e1", 0, 0, e1, ssss);
			eq.add("T", 0, 0, stress, ssss);
			eq.add("even", 0, 0, ssss);
			boot.add_equation(eq);
		}
		{
			qboot::Equation eq(boot, Odd);
			eq.add("unit", unit, eeee);
			eq.add("scalar", 1, 1, e, eeee);
			eq.add("e1", 1, 1, e1, eeee);
			eq.add("T", 1, 1, stress, eeee);
			eq.add("even", 1, 1, eeee);
			boot.add_equation(eq);
		}
		{
			qboot::Equation eq(boot, Odd);
			eq.add("scalar", 0, 0, s, eses);
			eq.add("odd+", eses);
			eq.add("odd-", eses);
			boot.add_equation(eq);
		}
		{
			qboot::Equation eq(boot, Odd);
			eq.add("unit", unit, eess);
			eq.add("scalar", 0, 0, s, esse);
			eq.add("scalar", 0, 1, e, eess);
			eq.add("e1", 0, 1, e1, eess);
			eq.add("T", 0, 1, stress, eess);
			eq.add("odd+", esse);
			eq.add("odd-", real(-1), esse);
			eq.add("even", 0, 1, eess);
			boot.add_equation(eq);
		}
		{
			qboot::Equation eq(boot, Even);
			eq.add("unit", unit, eess);
			eq.add("scalar", 0, 0, real(-1), s, esse);
			eq.add("scalar", 0, 1, e, eess);
			eq.add("e1", 0, 1, e1, eess);
			eq.add("T", 0, 1, stress, eess);
			eq.add("odd+", real(-1), esse);
			eq.add("odd-", esse);
			eq.add("even", 0, 1, eess);
			boot.add_equation(eq);
		}
		boot.finish();
		return boot;
	}
	double now()
	{
		return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
	}
}  // namespace

extern "C"
{
	int ref_words(int prec) { return int(nlimbs(prec)) + 2; }

	// decimal string -> number (include/qboot/mp/real.hpp:243-255, mpfr_set_str base 0/10, RNDN)
	int ref_from_str(int prec, const char* s, uint32_t* out)
	{
		set_prec(prec);
		try
		{
			real r{std::string_view(s)};
			put(r, prec, out);
		}
		catch (...)
		{
			return -1;
		}
		return 0;
	}
	// number -> the reference's ostream text at `digits` significant digits (include/qboot/mp/real.hpp:918-1030);
	// digits <= 0 selects the SDPB writer's 3 + int(prec * 0.302) (src/sdpb_input.cpp:32-35)
	int ref_to_str(int prec, const uint32_t* x, int digits, char* buf, int buflen)
	{
		set_prec(prec);
		std::ostringstream os;
		os.precision(digits > 0 ? digits : 3 + int(double(prec) * 0.302));
		os << mk(x, prec);
		auto s = os.str();
		if (int(s.size()) + 1 > buflen) return -int(s.size()) - 1;
		std::memcpy(buf, s.c_str(), s.size() + 1);
		return int(s.size());
	}
	// rational string ("37/10" or via parse "0.5181489") -> correctly rounded real (mpfr_set_q)
	int ref_from_rational(int prec, const char* s, uint32_t* out)
	{
		set_prec(prec);
		auto q = qboot::mp::parse(s);
		if (!q.has_value()) return -1;
		put(real(q.value()), prec, out);
		return 0;
	}

	// Elementwise MPFR primitives through the reference's mp::real operators, for pinning the device float.
	//   a, b, c : n numbers each (may be null when unused);  ia, ib : n int64 each (integer / rational operands)
	enum
	{
		OP_ADD = 0,
		OP_SUB = 1,
		OP_MUL = 2,
		OP_DIV = 3,
		OP_FMA = 4,      // a * b + c
		OP_SQRT = 5,
		OP_MUL_SI = 6,   // a * ia
		OP_ADD_SI = 7,   // a + ia
		OP_DIV_SI = 8,   // a / ia
		OP_MUL_Q = 9,    // a * (ia / ib)
		OP_DIV_Q = 10,   // a / (ia / ib)
		OP_ADD_Q = 11,   // a + (ia / ib)
		OP_SET_Q = 12,   // (ia / ib)
		OP_POW = 13,     // a ^ b
		OP_UI_POW = 14,  // ia ^ a
		OP_EXP = 15,
		OP_LOG = 16,
		OP_GAMMA_INC = 17,  // Gamma(a, b)
		OP_SI_SUB = 18,     // ia - a
		OP_SI_DIV = 19,     // ia / a
		OP_SUB_Q = 20,      // a - ia / ib
		OP_PI = 21,
		OP_SET_SI_2EXP = 22,  // ia * 2^ib
		OP_CMP = 23,          // out word 0 = cmp(a, b) + 1
		OP_ISINT = 24,        // out word 0 = isinteger(a)
		OP_FLOOR = 25,
	};
	int ref_op(int prec, int op, int n, const uint32_t* a, const uint32_t* b, const uint32_t* c, const int64_t* ia,
	           const int64_t* ib, uint32_t* out)
	{
		set_prec(prec);
		size_t W = size_t(ref_words(prec));
		for (int i = 0; i < n; ++i)
		{
			real x = a ? mk(a + W * size_t(i), prec) : real{};
			real y = b ? mk(b + W * size_t(i), prec) : real{};
			real z = c ? mk(c + W * size_t(i), prec) : real{};
			long p = ia ? long(ia[i]) : 0, q = ib ? long(ib[i]) : 1;
			real r;
			switch (op)
			{
			case OP_ADD: r = x + y; break;
			case OP_SUB: r = x - y; break;
			case OP_MUL: r = x * y; break;
			case OP_DIV: r = x / y; break;
			case OP_FMA: qboot::mp::fma(r, x, y, z); break;
			case OP_SQRT: r = qboot::mp::sqrt(x); break;
			case OP_MUL_SI: r = x * p; break;
			case OP_ADD_SI: r = x + p; break;
			case OP_DIV_SI: r = x / p; break;
			case OP_MUL_Q: r = x * rational(p, static_cast<unsigned long>(q)); break;
			case OP_DIV_Q: r = x / rational(p, static_cast<unsigned long>(q)); break;
			case OP_ADD_Q: r = x + rational(p, static_cast<unsigned long>(q)); break;
			case OP_SET_Q: r = real(rational(p, static_cast<unsigned long>(q))); break;
			case OP_POW: r = qboot::mp::pow(x, y); break;
			case OP_UI_POW: r = qboot::mp::pow(static_cast<unsigned long>(p), x); break;
			case OP_EXP: r = qboot::mp::exp(x); break;
			case OP_LOG: r = qboot::mp::log(x); break;
			case OP_GAMMA_INC: r = qboot::mp::gamma_inc(x, y); break;
			case OP_SI_SUB: r = p - x; break;
			case OP_SI_DIV: r = p / x; break;
			case OP_SUB_Q: r = x - rational(p, static_cast<unsigned long>(q)); break;
			case OP_PI: r = qboot::mp::const_pi(); break;
			case OP_SET_SI_2EXP: r = real(p, q); break;
			case OP_CMP:
			{
				std::memset(out + W * size_t(i), 0, W * 4);
				out[W * size_t(i)] = uint32_t((x < y) ? 0 : (x == y) ? 1 : 2);
				continue;
			}
			case OP_ISINT:
			{
				std::memset(out + W * size_t(i), 0, W * 4);
				out[W * size_t(i)] = qboot::mp::isinteger(x) ? 1u : 0u;
				continue;
			}
			case OP_FLOOR: r = qboot::mp::floor(x); break;
			default: return -1;
			}
			put(r, prec, out + W * size_t(i));
		}
		return 0;
	}

	// Context constants (src/context.cpp:47-61): rho = 3 - sqrt(8) and the rho->z matrix, row-major (lambda+1)^2.
	// The matrix is private; row j is recovered exactly as convert(e_j) (real_function.hpp:208-213).
	int ref_context_consts(int prec, uint32_t n_max, uint32_t lambda, const char* dim, uint32_t* rho, uint32_t* mat)
	{
		const auto& c = get_ctx(prec, n_max, lambda, dim);
		set_prec(prec);
		size_t W = size_t(ref_words(prec));
		put(c.rho(), prec, rho);
		for (uint32_t j = 0; j <= lambda; ++j)
		{
			alg::RealFunction<real> e(lambda);
			e.at(j) = 1;
			auto row = c.rho_to_z().convert(e);
			for (uint32_t i = 0; i <= lambda; ++i) put(row.at(i), prec, mat + W * (size_t(j) * (lambda + 1) + i));
		}
		return 0;
	}

	// _get_rec_coeffs (include/qboot/hor_formula.hpp:34-39): out = K polynomials x (deg+1) numbers, K=6,deg=3 for
	// spin 0 and K=8,deg=4 otherwise; coefficient of n^d at out[(i*(deg+1)+d)]
	int ref_rec_coeffs(int prec, const char* dim, const uint32_t* delta, uint32_t spin, const uint32_t* S,
	                   const uint32_t* P, uint32_t* out)
	{
		set_prec(prec);
		size_t W = size_t(ref_words(prec));
		Op op(mk(delta, prec), spin, eps_of(dim));
		auto ps = qboot::_get_rec_coeffs(op, mk(S, prec), mk(P, prec));
		uint32_t deg = spin == 0 ? 3 : 4;
		for (uint32_t i = 0; i < ps.size(); ++i)
			for (uint32_t d = 0; d <= deg; ++d)
			{
				real v = int32_t(d) <= ps[i].degree() ? ps[i].at(d) : real{};
				put(v, prec, out + W * (size_t(i) * (deg + 1) + d));
			}
		return int(ps.size());
	}
	// hBlock_shifted (src/hor_recursion.cpp:14-39): out = n_max+1 numbers
	int ref_hblock(int prec, uint32_t n_max, const char* dim, const uint32_t* delta, uint32_t spin, const uint32_t* S,
	               const uint32_t* P, uint32_t* out)
	{
		set_prec(prec);
		size_t W = size_t(ref_words(prec));
		Op op(mk(delta, prec), spin, eps_of(dim));
		auto b = qboot::hBlock_shifted(op, mk(S, prec), mk(P, prec), n_max);
		for (uint32_t n = 0; n <= n_max; ++n) put(b.at(n), prec, out + W * n);
		return 0;
	}
	// gBlock_real (src/hor_recursion.cpp:40-56): out = lambda+1 numbers (z-derivatives on the diagonal)
	int ref_gblock_real(int prec, uint32_t n_max, uint32_t lambda, const char* dim, const uint32_t* delta, uint32_t spin,
	                    const uint32_t* S, const uint32_t* P, uint32_t* out)
	{
		const auto& c = get_ctx(prec, n_max, lambda, dim);
		set_prec(prec);
		size_t W = size_t(ref_words(prec));
		Op op(mk(delta, prec), spin, c);
		auto g = qboot::gBlock_real(op, mk(S, prec), mk(P, prec), c);
		for (uint32_t k = 0; k <= lambda; ++k) put(g.at(k), prec, out + W * k);
		return 0;
	}
	// gBlock (src/hor_recursion.cpp:57-60): n tables, out = n * dim_M numbers in flatten() order
	// (dy-major, dx-minor, include/qboot/algebra/complex_function.hpp:72-81).  Returns dim_M.
	int ref_gblock(int prec, uint32_t n_max, uint32_t lambda, const char* dim, int n, const uint32_t* delta,
	               const uint32_t* spin, const uint32_t* S, const uint32_t* P, uint32_t* out)
	{
		const auto& c = get_ctx(prec, n_max, lambda, dim);
		set_prec(prec);
		size_t W = size_t(ref_words(prec));
		size_t dimM = alg::function_dimension(lambda, alg::FunctionSymmetry::Mixed);
		for (int t = 0; t < n; ++t)
		{
			Op op(mk(delta + W * size_t(t), prec), spin[t], c);
			auto g = qboot::gBlock(op, mk(S + W * size_t(t), prec), mk(P + W * size_t(t), prec), c);
			auto v = std::move(g).flatten();
			for (size_t i = 0; i < dimM; ++i) put(v[uint32_t(i)], prec, out + W * (size_t(t) * dimM + i));
		}
		return int(dimM);
	}
	// F/H block of one term (include/qboot/context.hpp:123-128): 2 * proj_sym(v^d * g); sym 0 = Even (H), 1 = Odd (F)
	int ref_fblock(int prec, uint32_t n_max, uint32_t lambda, const char* dim, const uint32_t* d, const uint32_t* delta,
	               uint32_t spin, const uint32_t* S, const uint32_t* P, int sym, uint32_t* out)
	{
		const auto& c = get_ctx(prec, n_max, lambda, dim);
		set_prec(prec);
		size_t W = size_t(ref_words(prec));
		Op op(mk(delta, prec), spin, c);
		auto g = qboot::gBlock(op, mk(S, prec), mk(P, prec), c);
		auto f = c.F_block(mk(d, prec), g, sym ? Odd : Even);
		auto v = std::move(f).flatten();
		for (uint32_t i = 0; i < v.size(); ++i) put(v[i], prec, out + W * i);
		return int(v.size());
	}
	// v^d table alone: F_block with g = 1 is not available, so use F_block on a unit "g" (at(0,0)=1) for both
	// parities and interleave -- instead expose mul(v_to_d, g) indirectly: callers compare F blocks. (kept minimal)

	// ConformalScale (src/conformal_scale.cpp:188-212, include/qboot/conformal_scale.hpp:23-117) for
	// GeneralPrimaryOperator(spin, num_poles, eps, lb[, ub]).  Outputs: D+1 sample points, D+1 sample scalings,
	// D+1 deltas, poles, and the bilinear bases (D/2+1 polynomials, coefficient-major, each padded to D/2+1).
	// Returns D (max_degree) or <0.
	int ref_conformal_scale(int prec, uint32_t n_max, uint32_t lambda, const char* dim, uint32_t spin, uint32_t num_poles,
	                        int include_odd, const uint32_t* lb, const uint32_t* ub /* may be null */, uint32_t* xs,
	                        uint32_t* scalings, uint32_t* deltas, uint32_t* bases, int bases_cap)
	{
		const auto& c = get_ctx(prec, n_max, lambda, dim);
		set_prec(prec);
		size_t W = size_t(ref_words(prec));
		std::optional<real> u;
		if (ub) u = mk(ub, prec);
		qboot::GeneralPrimaryOperator op(spin, num_poles, c.epsilon(), mk(lb, prec), u);
		qboot::ConformalScale sc(op, c, include_odd != 0);
		uint32_t D = sc.max_degree();
		auto sp = sc.sample_points();
		auto ss = sc.sample_scalings();
		for (uint32_t k = 0; k <= D; ++k)
		{
			put(sp[k], prec, xs + W * k);
			put(ss[k], prec, scalings + W * k);
			put(sc.get_delta(sp[k]), prec, deltas + W * k);
		}
		auto q = sc.bilinear_bases();
		uint32_t nb = q.size();
		if (int(nb * nb) > bases_cap) return -2;
		for (uint32_t i = 0; i < nb; ++i)
			for (uint32_t j = 0; j < nb; ++j)
			{
				real v = int32_t(j) <= q[i].degree() ? q[i].at(j) : real{};
				put(v, prec, bases + W * (size_t(i) * nb + j));
			}
		return int(D);
	}

	// Full PMP build through convert -> create_input -> write (src/bootstrap_equation.cpp:555-593,
	// src/polynomial_program.cpp:152-232, src/sdpb_input.cpp:138-199).
	//   model 0: single-correlator Ising (sample/main.cpp:32-59); model 1: mixed Ising (test/src/main.cpp:97-178)
	//   mode 0: FindContradiction("unit"); mode 1: ExtremalOPE(true, "T", "unit")
	//   outdir may be null/empty: then the SDPB directory is not written (write phase skipped).
	// times[0..2] = seconds of convert / create_input / write; counts[0] = #tables (gBlock memo entries),
	// counts[1] = #constraints, counts[2] = N (number of free variables before elimination)
	int ref_build_pmp(int prec, uint32_t n_max, uint32_t lambda, const char* dim, int model, int mode, int finite,
	                  const char* ds, const char* de, const char* de1, uint32_t numax, uint32_t maxspin,
	                  uint32_t threads, const char* outdir, double* times, int64_t* counts)
	{
		set_prec(prec);
		try
		{
			dict d;
			d["s"] = qboot::mp::parse(ds).value();
			d["e"] = qboot::mp::parse(de).value();
			if (de1 && *de1) d["e1"] = qboot::mp::parse(de1).value();
			qboot::Context c(n_max, lambda, rational(dim), threads);
			uint32_t base = c._total_memory();
			auto boot = model == 0 ? model_single(c, d, numax, maxspin, finite != 0)
			                       : model_mixed(c, d, numax, maxspin, finite != 0);
			double t0 = now();
			auto pmp = mode == 0 ? boot.convert(qboot::FindContradiction("unit"), threads)
			                     : boot.convert(qboot::ExtremalOPE(true, "T", "unit"), threads);
			double t1 = now();
			size_t dimM = alg::function_dimension(lambda, alg::FunctionSymmetry::Mixed);
			if (counts)
			{
				counts[0] = int64_t((c._total_memory() - base) / dimM);
				counts[2] = int64_t(pmp.num_of_variables());
			}
			auto input = std::move(pmp).create_input(threads);
			double t2 = now();
			if (counts) counts[1] = int64_t(input.num_constraints());
			if (outdir && *outdir) std::move(input).write(qboot::fs::path(outdir), threads);
			double t3 = now();
			if (times)
			{
				times[0] = t1 - t0;
				times[1] = t2 - t1;
				times[2] = t3 - t2;
			}
		}
		catch (const std::exception&)
		{
			return -1;
		}
		return 0;
	}

	// CPU timing of the block-table hot path alone: n tables through qboot::gBlock on `threads` host threads
	// (the same fan-out the reference uses, include/qboot/task_queue.hpp:33-60).  Returns seconds.
	double ref_time_gblock(int prec, uint32_t n_max, uint32_t lambda, const char* dim, int n, const uint32_t* delta,
	                       const uint32_t* spin, const uint32_t* S, const uint32_t* P, uint32_t threads)
	{
		const auto& c = get_ctx(prec, n_max, lambda, dim);
		set_prec(prec);
		size_t W = size_t(ref_words(prec));
		std::vector<std::function<uint32_t()>> tasks;
		for (int t = 0; t < n; ++t)
			tasks.emplace_back([&, t] {
				Op op(mk(delta + W * size_t(t), prec), spin[t], c);
				auto g = qboot::gBlock(op, mk(S + W * size_t(t), prec), mk(P + W * size_t(t), prec), c);
				return g.size();
			});
		double t0 = now();
		auto r = qboot::_parallel_evaluate(tasks, threads);
		double t1 = now();
		return r.size() == size_t(n) ? t1 - t0 : -1.0;
	}
}
