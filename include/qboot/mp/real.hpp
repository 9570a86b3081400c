// qboot_b200 façade -- qboot::mp::real on the engine's own number format.
//
// Same name and call surface as the reference's MPFR wrapper (include/qboot/mp/real.hpp:87-612 and the free functions
// at :688-855), so model code written against qboot compiles unchanged.  A `real` here is one number record of the
// engine (include/qboot_b200.h): L = ceil(global_prec / 32) limbs, held inline -- the very bytes that are uploaded to
// and read back from the GPU, no conversion in between.  Arithmetic is forwarded to qb::host (hostops.hpp): every
// operation is rounded once to nearest-even at global_prec bits, as mpfr_* with MPFR_RNDN rounds it.
//
// Differences from the reference, all documented in DESIGN.md:
//   * mp::global_prec must be set before the first real is made and must name a precision with compiled kernels
//     (limb counts 16/19/25/32/38/48, i.e. 481-512, 577-608, 769-800, 993-1024, 1185-1216, 1505-1536 bits);
//   * only MPFR_RNDN is implemented (the reference's drivers never use another mode).
#ifndef QBOOT_B200_MP_REAL_HPP_
#define QBOOT_B200_MP_REAL_HPP_

#include <cstdint>
#include <cstring>
#include <ios>
#include <istream>
#include <optional>
#include <ostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <string_view>
#include <type_traits>
#include <utility>

#include "qboot/b200/hostops.hpp"
#include "qboot/mp/rational.hpp"

// drivers written for the reference assign `qboot::mp::global_rnd = MPFR_RNDN;` (sample/main.cpp:64)
#ifndef MPFR_VERSION
using mpfr_prec_t = long;
enum mpfr_rnd_t
{
	MPFR_RNDN = 0,
	MPFR_RNDZ,
	MPFR_RNDU,
	MPFR_RNDD,
	MPFR_RNDA
};
#endif

namespace qboot::mp
{
	extern mpfr_prec_t global_prec;  // defined in qboot_b200/host/facade.cpp (default 1000, src/mp/real.cpp:5)
	extern mpfr_rnd_t global_rnd;
	inline int _prec() { return int(global_prec); }

	template <class T>
	inline constexpr bool _is_machine_int = std::is_integral_v<T> && !std::is_same_v<T, bool>;

	class real
	{
		uint32_t w_[qb::host::kMaxWords];
		static size_t bytes() { return 4 * size_t((_prec() + 31) / 32 + 2); }

	public:
		// raw record access (what the planner uploads)
		[[nodiscard]] const uint32_t* _words() const { return w_; }
		[[nodiscard]] uint32_t* _words() { return w_; }
		void _reset() && { std::memset(w_, 0, sizeof(w_)); }

		real() { std::memset(w_, 0, sizeof(w_)); }
		real(const real& o) = default;
		real& operator=(const real& o) & = default;
		template <class T, class = std::enable_if_t<_is_machine_int<T>>>
		explicit real(T v)
		{
			std::memset(w_, 0, sizeof(w_));
			if constexpr (std::is_signed_v<T>)
				qb::host::h_set_si(w_, int64_t(v), _prec());
			else
				qb::host::h_set_ui(w_, uint64_t(v), _prec());
		}
		explicit real(double v)
		{
			std::memset(w_, 0, sizeof(w_));
			qb::host::h_set_d(w_, v, _prec());
		}
		explicit real(const rational& q)
		{
			std::memset(w_, 0, sizeof(w_));
			qb::host::h_set_q(w_, q.num(), q.den(), _prec());
		}
		// v * 2^exp
		template <class T, class = std::enable_if_t<_is_machine_int<T>>>
		real(T v, long exp)
		{
			std::memset(w_, 0, sizeof(w_));
			qb::host::h_set_si_2exp(w_, int64_t(v), exp, _prec());
		}
		explicit real(std::string_view s)
		{
			std::memset(w_, 0, sizeof(w_));
			if (!qb::host::h_from_string(w_, s, _prec()))
				throw std::runtime_error("in qboot::mp::real(string_view):\n  invalid input format " + std::string(s));
		}
		explicit real(const char* s) : real(std::string_view(s)) {}
		template <class T, class = std::enable_if_t<_is_machine_int<T>>>
		real& operator=(T v) &
		{
			return *this = real(v);
		}
		real& operator=(const rational& q) & { return *this = real(q); }

		[[nodiscard]] real clone() const { return *this; }
		void swap(real& o) & { std::swap(*this, o); }
		[[nodiscard]] bool iszero() const { return (w_[0] & 3u) == 0u; }
		[[nodiscard]] bool isnan() const { return (w_[0] & 3u) == 3u; }
		[[nodiscard]] bool isinf() const { return (w_[0] & 3u) == 2u; }
		[[nodiscard]] bool isinteger() const { return qb::host::h_is_integer(w_, _prec()); }
		void negate() & { qb::host::h_neg(w_, w_, _prec()); }
		[[nodiscard]] real norm() const { return *this * *this; }
		[[nodiscard]] real eval([[maybe_unused]] const real& x) const { return *this; }
		// text in the stream's default notation at 6 / `precision` significant digits
		[[nodiscard]] std::string str() const;
		[[nodiscard]] std::string str(int32_t precision) const;
		explicit operator double() const { return qb::host::h_get_d(w_, _prec()); }

		real operator+() const { return *this; }
		real operator-() const
		{
			real r = *this;
			r.negate();
			return r;
		}
		// ---- real (op) real ----
		friend real operator+(const real& a, const real& b)
		{
			real r;
			qb::host::h_add(r.w_, a.w_, b.w_, _prec());
			return r;
		}
		friend real operator-(const real& a, const real& b)
		{
			real r;
			qb::host::h_sub(r.w_, a.w_, b.w_, _prec());
			return r;
		}
		friend real operator*(const real& a, const real& b)
		{
			real r;
			qb::host::h_mul(r.w_, a.w_, b.w_, _prec());
			return r;
		}
		friend real operator/(const real& a, const real& b)
		{
			real r;
			qb::host::h_div(r.w_, a.w_, b.w_, _prec());
			return r;
		}
		// ---- real (op) machine integer: mpfr_*_si / mpfr_*_ui ----
		template <class T, class = std::enable_if_t<_is_machine_int<T>>>
		friend real operator+(const real& a, T v)
		{
			real r;
			qb::host::h_add_si(r.w_, a.w_, int64_t(v), _prec());
			return r;
		}
		template <class T, class = std::enable_if_t<_is_machine_int<T>>>
		friend real operator+(T v, const real& a)
		{
			return a + v;
		}
		template <class T, class = std::enable_if_t<_is_machine_int<T>>>
		friend real operator-(const real& a, T v)
		{
			real r;
			qb::host::h_add_si(r.w_, a.w_, -int64_t(v), _prec());
			return r;
		}
		template <class T, class = std::enable_if_t<_is_machine_int<T>>>
		friend real operator-(T v, const real& a)
		{
			real r;
			qb::host::h_si_sub(r.w_, int64_t(v), a.w_, _prec());
			return r;
		}
		template <class T, class = std::enable_if_t<_is_machine_int<T>>>
		friend real operator*(const real& a, T v)
		{
			real r;
			qb::host::h_mul_si(r.w_, a.w_, int64_t(v), _prec());
			return r;
		}
		template <class T, class = std::enable_if_t<_is_machine_int<T>>>
		friend real operator*(T v, const real& a)
		{
			return a * v;
		}
		template <class T, class = std::enable_if_t<_is_machine_int<T>>>
		friend real operator/(const real& a, T v)
		{
			real r;
			qb::host::h_div_si(r.w_, a.w_, int64_t(v), _prec());
			return r;
		}
		template <class T, class = std::enable_if_t<_is_machine_int<T>>>
		friend real operator/(T v, const real& a)
		{
			real r;
			qb::host::h_si_div(r.w_, int64_t(v), a.w_, _prec());
			return r;
		}
		// ---- real (op) rational: mpfr_*_q ----
		friend real operator+(const real& a, const rational& q)
		{
			real r;
			qb::host::h_add_q(r.w_, a.w_, q.num(), q.den(), _prec());
			return r;
		}
		friend real operator+(const rational& q, const real& a) { return a + q; }
		friend real operator-(const real& a, const rational& q)
		{
			real r;
			qb::host::h_sub_q(r.w_, a.w_, q.num(), q.den(), _prec());
			return r;
		}
		friend real operator-(const rational& q, const real& a)
		{
			real r = a - q;
			r.negate();
			return r;
		}
		friend real operator*(const real& a, const rational& q)
		{
			real r;
			qb::host::h_mul_q(r.w_, a.w_, q.num(), q.den(), _prec());
			return r;
		}
		friend real operator*(const rational& q, const real& a) { return a * q; }
		friend real operator/(const real& a, const rational& q)
		{
			real r;
			qb::host::h_div_q(r.w_, a.w_, q.num(), q.den(), _prec());
			return r;
		}
		friend real operator/(const rational& q, const real& a) { return real(q) / a; }

		template <class T>
		real& operator+=(const T& o) &
		{
			return *this = *this + o;
		}
		template <class T>
		real& operator-=(const T& o) &
		{
			return *this = *this - o;
		}
		template <class T>
		real& operator*=(const T& o) &
		{
			return *this = *this * o;
		}
		template <class T>
		real& operator/=(const T& o) &
		{
			return *this = *this / o;
		}

		// ---- comparisons (nan compares unordered) ----
		friend int _cmp(const real& a, const real& b) { return qb::host::h_cmp(a.w_, b.w_, _prec()); }
		template <class T, class = std::enable_if_t<_is_machine_int<T>>>
		friend int _cmp(const real& a, T v)
		{
			return qb::host::h_cmp_si(a.w_, int64_t(v), _prec());
		}
		friend int _cmp(const real& a, const rational& q) { return qb::host::h_cmp_q(a.w_, q.num(), q.den(), _prec()); }
		[[nodiscard]] bool _unordered() const { return isnan(); }
	};

#define QBOOT_B200_REAL_RELOP(op)                                                                   \
	inline bool operator op(const real& a, const real& b)                                           \
	{                                                                                               \
		return !a.isnan() && !b.isnan() && _cmp(a, b) op 0;                                         \
	}                                                                                               \
	template <class T, class = std::enable_if_t<_is_machine_int<T> || std::is_same_v<T, rational>>> \
	inline bool operator op(const real& a, const T& b)                                              \
	{                                                                                               \
		return !a.isnan() && _cmp(a, b) op 0;                                                       \
	}                                                                                               \
	template <class T, class = std::enable_if_t<_is_machine_int<T> || std::is_same_v<T, rational>>> \
	inline bool operator op(const T& a, const real& b)                                              \
	{                                                                                               \
		return !b.isnan() && 0 op _cmp(b, a);                                                       \
	}
	QBOOT_B200_REAL_RELOP(<)
	QBOOT_B200_REAL_RELOP(>)
	QBOOT_B200_REAL_RELOP(<=)
	QBOOT_B200_REAL_RELOP(>=)
	QBOOT_B200_REAL_RELOP(==)
#undef QBOOT_B200_REAL_RELOP
	inline bool operator!=(const real& a, const real& b) { return !(a == b); }
	template <class T, class = std::enable_if_t<_is_machine_int<T> || std::is_same_v<T, rational>>>
	inline bool operator!=(const real& a, const T& b)
	{
		return !(a == b);
	}
	template <class T, class = std::enable_if_t<_is_machine_int<T> || std::is_same_v<T, rational>>>
	inline bool operator!=(const T& a, const real& b)
	{
		return !(a == b);
	}

	// ---- free functions (include/qboot/mp/real.hpp:665-855) ---------------------------------------------------------
	inline real mul(const real& a, const real& b) { return a * b; }
	inline real mul_scalar(const real& a, const real& b) { return a * b; }
	template <class T, class = std::enable_if_t<_is_machine_int<T>>>
	inline real mul_scalar(T c, const real& x)
	{
		return c * x;
	}
	inline bool iszero(const real& r) { return r.iszero(); }
	inline bool isinteger(const real& r) { return r.isinteger(); }
	inline int sgn(const real& r) { return r.iszero() || r.isnan() ? 0 : ((r._words()[0] >> 31) ? -1 : 1); }
	inline int cmpabs(const real& a, const real& b) { return qb::host::h_cmpabs(a._words(), b._words(), _prec()); }
	inline real abs(const real& r)
	{
		real a = r;
		qb::host::h_abs(a._words(), a._words(), _prec());
		return a;
	}
	inline real zero(int n = 1)
	{
		real r;
		if (n < 0) r.negate();
		return r;
	}
	inline real inf(int n = 1) { return real(n < 0 ? "-inf" : "inf"); }
	inline real nan() { return real("nan"); }
#define QBOOT_B200_REAL_UNARY(name, impl)         \
	inline real name(const real& x)               \
	{                                             \
		real r;                                   \
		qb::host::impl(r._words(), x._words(), _prec()); \
		return r;                                 \
	}                                             \
	inline void name(real& ans, const real& x) { qb::host::impl(ans._words(), x._words(), _prec()); }
	QBOOT_B200_REAL_UNARY(sqrt, h_sqrt)
	QBOOT_B200_REAL_UNARY(exp, h_exp)
	QBOOT_B200_REAL_UNARY(log, h_log)
	QBOOT_B200_REAL_UNARY(floor, h_floor)
	QBOOT_B200_REAL_UNARY(ceil, h_ceil)
	QBOOT_B200_REAL_UNARY(round, h_round)
#undef QBOOT_B200_REAL_UNARY
	inline real sqrt(unsigned long v) { return sqrt(real(v)); }
	inline real const_pi() noexcept
	{
		real r;
		qb::host::h_const_pi(r._words(), _prec());
		return r;
	}
	inline real const_log2()
	{
		real r;
		qb::host::h_const_log2(r._words(), _prec());
		return r;
	}
	// ans = r1 * r2 + r3, rounded once (mpfr_fma); the others likewise
	inline void fma(real& ans, const real& r1, const real& r2, const real& r3)
	{
		qb::host::h_fma(ans._words(), r1._words(), r2._words(), r3._words(), _prec());
	}
	inline void fms(real& ans, const real& r1, const real& r2, const real& r3)
	{
		qb::host::h_fms(ans._words(), r1._words(), r2._words(), r3._words(), _prec());
	}
	inline real fma(const real& r1, const real& r2, const real& r3)
	{
		real r;
		fma(r, r1, r2, r3);
		return r;
	}
	inline real fms(const real& r1, const real& r2, const real& r3)
	{
		real r;
		fms(r, r1, r2, r3);
		return r;
	}
	inline real gamma_inc(const real& a, const real& x)
	{
		real r;
		qb::host::h_gamma_inc(r._words(), a._words(), x._words(), _prec());
		return r;
	}
	inline void gamma_inc(real& ans, const real& a, const real& x) { ans = gamma_inc(a, x); }
	inline real pow(const real& x, const real& y)
	{
		real r;
		qb::host::h_pow(r._words(), x._words(), y._words(), _prec());
		return r;
	}
	inline void pow(real& ans, const real& x, const real& y) { ans = pow(x, y); }
	inline real pow(const real& x, unsigned long n)
	{
		real r;
		qb::host::h_pow_si(r._words(), x._words(), int64_t(n), _prec());
		return r;
	}
	inline real pow(const real& x, long n)
	{
		real r;
		qb::host::h_pow_si(r._words(), x._words(), int64_t(n), _prec());
		return r;
	}
	inline real pow(const real& x, unsigned int n) { return pow(x, static_cast<unsigned long>(n)); }
	inline real pow(const real& x, int n) { return pow(x, long(n)); }
	inline real pow(unsigned long base, const real& y)
	{
		real r;
		qb::host::h_ui_pow(r._words(), base, y._words(), _prec());
		return r;
	}
	inline real pochhammer(const real& x, unsigned long n)
	{
		real r(1);
		for (unsigned long i = 0; i < n; ++i) r *= x + i;
		return r;
	}

	// ---- text (include/qboot/mp/real.hpp:860-1133): default-float / fixed / scientific like the reference's operator<< ----
	inline std::string _format(const real& x, std::ios_base::fmtflags flags, std::streamsize precision)
	{
		if (x.isnan()) return "@NaN@";
		const bool negative = (x._words()[0] >> 31) != 0;
		std::string sign = negative ? "-" : (flags & std::ios_base::showpos) ? "+" : "";
		if (x.isinf()) return sign + "@Inf@";
		if (x.iszero()) return sign + "0";
		const auto ff = flags & std::ios_base::floatfield;
		const bool fixed = ff == std::ios_base::fixed, sci = ff == std::ios_base::scientific;
		size_t prec = precision < 0 ? 6 : size_t(precision);
		if (!fixed && !sci && prec == 0) prec = 1;
		if (sci) ++prec;
		int64_t e = 0;
		size_t extra = 0;
		if (fixed)
		{
			int64_t e0 = 0;
			qb::host::h_get_digits(x._words(), _prec(), 1, &e0);
			extra = size_t(e0 < 0 ? 0 : e0) + 3;
		}
		std::string s = qb::host::h_get_digits(x._words(), _prec(), prec + extra, &e);
		auto scientific = [&](std::string&& t) {
			if (t.length() > 1) t.insert(1, 1, '.');
			int64_t ex = e - 1;
			t += (flags & std::ios_base::uppercase) ? 'E' : 'e';
			t += ex >= 0 ? '+' : '-';
			if (ex < 0) ex = -ex;
			if (ex <= 9) t += '0';
			return t + std::to_string(ex);
		};
		if (sci) return sign + scientific(std::move(s));
		if (fixed)
		{
			if (prec == 0) return sign + (e <= 0 ? std::string("0") : s);
			if (e <= 0)
			{
				s.insert(0, size_t(2 - e), '0');
				s[1] = '.';
				if (s.length() > 2 + prec) s.erase(2 + prec);
			}
			else
			{
				s.insert(size_t(e), 1, '.');
				if (s.length() > size_t(e) + 1 + prec) s.erase(size_t(e) + 1 + prec);
			}
			return sign + s;
		}
		if (-3 <= e && e <= int64_t(prec))
		{
			if (e <= 0)
			{
				s.insert(0, size_t(2 - e), '0');
				s[1] = '.';
			}
			else
				s.insert(size_t(e), 1, '.');
			while (s.back() == '0') s.pop_back();
			if (s.back() == '.') s.pop_back();
			return sign + s;
		}
		while (s.back() == '0') s.pop_back();
		return sign + scientific(std::move(s));
	}
	inline std::string real::str() const { return _format(*this, std::ios_base::fmtflags{}, 6); }
	inline std::string real::str(int32_t precision) const { return _format(*this, std::ios_base::fmtflags{}, precision); }
	template <class Char, class Traits>
	std::basic_ostream<Char, Traits>& operator<<(std::basic_ostream<Char, Traits>& os, const real& x)
	{
		std::string t = _format(x, os.flags(), os.precision());
		// sign-aware padding as the stream asks for it
		const auto w = os.width();
		os.width(0);
		const size_t len = t.size();
		const bool has_sign = !t.empty() && (t[0] == '-' || t[0] == '+');
		std::basic_string<Char, Traits> o;
		auto put = [&](std::string_view v) {
			for (char c : v) o.push_back(os.widen(c));
		};
		const size_t pad = w > 0 && size_t(w) > len ? size_t(w) - len : 0;
		const auto adj = os.flags() & std::ios_base::adjustfield;
		if (adj == std::ios_base::left)
		{
			put(t);
			o.append(pad, os.fill());
		}
		else if (adj == std::ios_base::internal && has_sign)
		{
			put(std::string_view(t).substr(0, 1));
			o.append(pad, os.fill());
			put(std::string_view(t).substr(1));
		}
		else
		{
			o.append(pad, os.fill());
			put(t);
		}
		return os << o;
	}
	template <class Char, class Traits>
	std::basic_istream<Char, Traits>& operator>>(std::basic_istream<Char, Traits>& in, real& r)
	{
		typename std::basic_istream<Char, Traits>::sentry s(in, false);
		if (!s) return in;
		// [+-]? digits* [.]? digits* ( [eE] [+-]? digits+ )?
		std::string num;
		bool digits = false, point = false, exp = false, edigits = false;
		for (;;)
		{
			auto pc = in.peek();
			if (pc == Traits::eof()) break;
			char c = in.narrow(Traits::to_char_type(pc), '\0');
			bool take = false;
			if (c >= '0' && c <= '9')
			{
				take = true;
				(exp ? edigits : digits) = true;
			}
			else if (c == '.' && !point && !exp)
				take = point = true;
			else if ((c == 'e' || c == 'E') && !exp && digits)
				take = exp = true;
			else if ((c == '+' || c == '-') && (num.empty() || num.back() == 'e' || num.back() == 'E'))
				take = true;
			if (!take) break;
			num.push_back(c);
			in.get();
		}
		if (!digits || (exp && !edigits) || !qb::host::h_from_string(r._words(), num, _prec()))
		{
			in.setstate(std::ios_base::failbit);
			r = real();
		}
		return in;
	}
}  // namespace qboot::mp

#endif  // QBOOT_B200_MP_REAL_HPP_
