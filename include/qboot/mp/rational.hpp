// qboot_b200 façade -- exact rationals for the model description (dimension, epsilon, external scaling dimensions).
//
// Same name and call surface as the reference's qboot::mp::rational (include/qboot/mp/rational.hpp), which wraps GMP's
// mpq_t.  On this engine rationals only ever describe a model (dim = 3, epsilon = 1/2, "0.5181489", pole positions) and
// are handed to the device as (int64 numerator, uint32 denominator), so the value range is 64-bit; an operation whose
// exact result leaves that range throws std::overflow_error instead of silently growing (documented limit).
#ifndef QBOOT_B200_MP_RATIONAL_HPP_
#define QBOOT_B200_MP_RATIONAL_HPP_

#include <cstdint>
#include <optional>
#include <ostream>
#include <stdexcept>
#include <string>
#include <string_view>
#include <type_traits>

namespace qboot::mp
{
	class rational
	{
		int64_t n_ = 0, d_ = 1;  // lowest terms, d_ > 0
		using wide = __int128;
		static int64_t narrow(wide v)
		{
			if (v > wide(INT64_MAX) || v < wide(INT64_MIN)) throw std::overflow_error("qboot::mp::rational: value exceeds 64 bits");
			return int64_t(v);
		}
		static wide gcd(wide a, wide b)
		{
			if (a < 0) a = -a;
			if (b < 0) b = -b;
			while (b)
			{
				wide t = a % b;
				a = b;
				b = t;
			}
			return a;
		}
		static rational make(wide n, wide d)
		{
			if (d == 0) throw std::domain_error("qboot::mp::rational: zero denominator");
			if (d < 0)
			{
				n = -n;
				d = -d;
			}
			wide g = gcd(n, d);
			if (g > 1)
			{
				n /= g;
				d /= g;
			}
			rational r;
			r.n_ = narrow(n);
			r.d_ = narrow(d);
			return r;
		}

	public:
		void _reset() &&
		{
			n_ = 0;
			d_ = 1;
		}
		rational() = default;
		template <class T, class = std::enable_if_t<std::is_integral_v<T>>>
		rational(T v) : n_(int64_t(v))  // NOLINT: integers convert implicitly, as with mpq
		{
		}
		template <class T, class U, class = std::enable_if_t<std::is_integral_v<T> && std::is_integral_v<U>>>
		rational(T num, U den)
		{
			*this = make(wide(num), wide(den));
		}
		// "123", "-17/31" (the forms mpq_set_str accepts); use parse() for decimals
		explicit rational(std::string_view s)
		{
			auto v = _parse(s);
			if (!v) throw std::runtime_error("in qboot::mp::rational(string_view):\n  invalid input format " + std::string(s));
			*this = *v;
		}
		static std::optional<rational> _parse(std::string_view s)
		{
			auto integer = [](std::string_view t, bool allow_sign) -> std::optional<wide> {
				size_t i = 0;
				bool neg = false;
				if (allow_sign && i < t.size() && (t[i] == '-' || t[i] == '+')) neg = t[i++] == '-';
				if (i >= t.size()) return {};
				wide v = 0;
				for (; i < t.size(); ++i)
				{
					if (t[i] < '0' || t[i] > '9') return {};
					v = v * 10 + (t[i] - '0');
					if (v > (wide(1) << 100)) return {};
				}
				return neg ? -v : v;
			};
			auto slash = s.find('/');
			auto num = integer(s.substr(0, slash), true);
			if (!num) return {};
			wide den = 1;
			if (slash != std::string_view::npos)
			{
				auto dv = integer(s.substr(slash + 1), false);
				if (!dv || *dv == 0) return {};
				den = *dv;
			}
			return make(*num, den);
		}
		[[nodiscard]] int64_t num() const { return n_; }
		[[nodiscard]] int64_t den() const { return d_; }
		[[nodiscard]] rational clone() const { return *this; }
		[[nodiscard]] bool iszero() const { return n_ == 0; }
		[[nodiscard]] bool isinteger() const { return d_ == 1; }
		void negate() & { n_ = narrow(-wide(n_)); }
		void invert() & { *this = make(d_, n_); }
		void swap(rational& o) &
		{
			std::swap(n_, o.n_);
			std::swap(d_, o.d_);
		}
		template <class T>
		[[nodiscard]] rational eval([[maybe_unused]] const T& x) const
		{
			return *this;
		}
		[[nodiscard]] std::string str() const { return str('/'); }
		// '/' replaced by slash, '-' by minus: rational(-17, 31).str('#', '%') == "%17#31"
		[[nodiscard]] std::string str(char slash, char minus = '-') const
		{
			std::string s;
			if (n_ < 0)
			{
				s += minus;
				s += std::to_string(-wide(n_) > wide(INT64_MAX) ? uint64_t(INT64_MAX) + 1 : uint64_t(-n_));
			}
			else
				s += std::to_string(n_);
			if (d_ != 1)
			{
				s += slash;
				s += std::to_string(d_);
			}
			return s;
		}
		rational operator+() const { return *this; }
		rational operator-() const { return make(-wide(n_), d_); }
		friend rational operator+(const rational& a, const rational& b) { return make(wide(a.n_) * b.d_ + wide(b.n_) * a.d_, wide(a.d_) * b.d_); }
		friend rational operator-(const rational& a, const rational& b) { return make(wide(a.n_) * b.d_ - wide(b.n_) * a.d_, wide(a.d_) * b.d_); }
		friend rational operator*(const rational& a, const rational& b) { return make(wide(a.n_) * b.n_, wide(a.d_) * b.d_); }
		friend rational operator/(const rational& a, const rational& b) { return make(wide(a.n_) * b.d_, wide(a.d_) * b.n_); }
		rational& operator+=(const rational& o) & { return *this = *this + o; }
		rational& operator-=(const rational& o) & { return *this = *this - o; }
		rational& operator*=(const rational& o) & { return *this = *this * o; }
		rational& operator/=(const rational& o) & { return *this = *this / o; }
		friend int _cmp(const rational& a, const rational& b)
		{
			wide l = wide(a.n_) * b.d_, r = wide(b.n_) * a.d_;
			return l < r ? -1 : l > r ? 1 : 0;
		}
		friend bool operator==(const rational& a, const rational& b) { return a.n_ == b.n_ && a.d_ == b.d_; }
		friend bool operator!=(const rational& a, const rational& b) { return !(a == b); }
		friend bool operator<(const rational& a, const rational& b) { return _cmp(a, b) < 0; }
		friend bool operator>(const rational& a, const rational& b) { return _cmp(a, b) > 0; }
		friend bool operator<=(const rational& a, const rational& b) { return _cmp(a, b) <= 0; }
		friend bool operator>=(const rational& a, const rational& b) { return _cmp(a, b) >= 0; }
		template <class Char, class Traits>
		friend std::basic_ostream<Char, Traits>& operator<<(std::basic_ostream<Char, Traits>& s, const rational& q)
		{
			return s << q.str().c_str();
		}
	};
	inline int sgn(const rational& q) { return q.num() < 0 ? -1 : q.num() > 0 ? 1 : 0; }
	inline bool iszero(const rational& q) { return q.iszero(); }
	inline rational abs(const rational& q) { return q.num() < 0 ? -q : q; }
	inline rational pow(const rational& q, unsigned long e)
	{
		rational r(1), b = q;
		while (e)
		{
			if (e & 1u) r *= b;
			e >>= 1;
			if (e) b *= b;
		}
		return r;
	}
	// (q)_n = q (q + 1) ... (q + n - 1)
	inline rational pochhammer(const rational& q, unsigned long n)
	{
		rational r(1);
		for (unsigned long i = 0; i < n; ++i) r *= q + rational(int64_t(i));
		return r;
	}

	// decimal / exponent / fraction notation without spaces -> exact rational:
	//   "123", "123/456", "1.25" (= 5/4), "3.2e-8", "-5e3";  "12 / 59" or a trailing space fails
	inline std::optional<rational> parse(std::string_view s)
	{
		if (s.empty()) return {};
		if (s.find_first_not_of("+-0123456789/") == std::string_view::npos) return rational::_parse(s);
		size_t i = 0;
		bool neg = false;
		if (s[i] == '+' || s[i] == '-') neg = s[i++] == '-';
		__int128 mant = 0;
		int64_t scale = 0;  // value = mant * 10^scale
		bool digits = false, point = false;
		for (; i < s.size() && s[i] != 'e' && s[i] != 'E'; ++i)
		{
			if (s[i] == '.')
			{
				if (point) return {};
				point = true;
				continue;
			}
			if (s[i] < '0' || s[i] > '9') return {};
			mant = mant * 10 + (s[i] - '0');
			if (mant > (__int128(1) << 100)) return {};
			digits = true;
			if (point) --scale;
		}
		if (!digits) return {};
		if (i < s.size())
		{
			std::string_view e = s.substr(i + 1);
			if (e.empty() || e.find_first_not_of("+-0123456789") != std::string_view::npos) return {};
			auto ev = rational::_parse(e);
			if (!ev || !ev->isinteger() || ev->num() > 400 || ev->num() < -400) return {};
			scale += ev->num();
		}
		try
		{
			rational r = rational(int64_t(1));
			rational m;
			{
				// mant may exceed 64 bits only for absurd inputs; make() reports that
				rational ten(10);
				rational p = pow(ten, static_cast<unsigned long>(scale < 0 ? -scale : scale));
				// build mant exactly through two 64-bit halves when needed
				__int128 hi = mant >> 62, lo = mant & ((__int128(1) << 62) - 1);
				m = rational(int64_t(lo));
				if (hi) m += rational(int64_t(hi)) * rational(int64_t(1) << 62);
				r = scale < 0 ? m / p : m * p;
			}
			return neg ? -r : r;
		}
		catch (const std::overflow_error&)
		{
			return {};
		}
	}
}  // namespace qboot::mp

#endif  // QBOOT_B200_MP_RATIONAL_HPP_
