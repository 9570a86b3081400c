// qboot_b200 host library -- arbitrary-size naturals and a working-precision binary float, host only.
//
// The reference gets its transcendental functions, decimal input and decimal output from MPFR (mpfr_exp, mpfr_log,
// mpfr_pow, mpfr_const_pi, mpfr_gamma_inc, mpfr_set_str, mpfr_get_str; include/qboot/mp/real.hpp:243-255,688-804,
// 987-1030), all of them correctly rounded.  MPFR is third-party and not part of this engine, so the host side
// restates them here on plain integer arithmetic: a function is evaluated at a working precision with a known error
// bound and rounded to the target precision only when the bound proves the rounding (Ziv's strategy), which yields
// the same bits as any other correctly rounding implementation.  These routines run once per constraint (scale
// factors, sample points) or per printed number, never per table entry.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace qb::host
{
	using u64 = uint64_t;
	using u128 = unsigned __int128;

	// natural number, little-endian 64-bit limbs, no leading zero limbs (zero = empty)
	struct Nat
	{
		std::vector<u64> d;
		Nat() = default;
		explicit Nat(u64 v)
		{
			if (v) d.push_back(v);
		}
		bool zero() const { return d.empty(); }
		void trim()
		{
			while (!d.empty() && d.back() == 0) d.pop_back();
		}
		int64_t bits() const { return d.empty() ? 0 : int64_t(d.size()) * 64 - __builtin_clzll(d.back()); }
		bool bit(int64_t i) const
		{
			size_t w = size_t(i >> 6);
			return i >= 0 && w < d.size() && ((d[w] >> (i & 63)) & 1u);
		}
		// any bit set strictly below position i
		bool any_below(int64_t i) const
		{
			if (i <= 0) return false;
			size_t w = size_t(i >> 6);
			for (size_t k = 0; k < w && k < d.size(); ++k)
				if (d[k]) return true;
			if (w < d.size() && (i & 63) && (d[w] & ((u64(1) << (i & 63)) - 1))) return true;
			return false;
		}
		u64 low64() const { return d.empty() ? 0 : d[0]; }
	};

	inline int cmp(const Nat& a, const Nat& b)
	{
		if (a.d.size() != b.d.size()) return a.d.size() < b.d.size() ? -1 : 1;
		for (size_t i = a.d.size(); i-- > 0;)
			if (a.d[i] != b.d[i]) return a.d[i] < b.d[i] ? -1 : 1;
		return 0;
	}
	inline Nat add(const Nat& a, const Nat& b)
	{
		const Nat& x = a.d.size() >= b.d.size() ? a : b;
		const Nat& y = a.d.size() >= b.d.size() ? b : a;
		Nat r;
		r.d.resize(x.d.size() + 1);
		u64 c = 0;
		for (size_t i = 0; i < x.d.size(); ++i)
		{
			u128 s = u128(x.d[i]) + (i < y.d.size() ? y.d[i] : 0) + c;
			r.d[i] = u64(s);
			c = u64(s >> 64);
		}
		r.d[x.d.size()] = c;
		r.trim();
		return r;
	}
	// a - b, a >= b
	inline Nat sub(const Nat& a, const Nat& b)
	{
		Nat r;
		r.d.resize(a.d.size());
		u64 br = 0;
		for (size_t i = 0; i < a.d.size(); ++i)
		{
			u64 y = i < b.d.size() ? b.d[i] : 0;
			u128 t = u128(a.d[i]) - y - br;
			r.d[i] = u64(t);
			br = (t >> 64) ? 1 : 0;
		}
		r.trim();
		return r;
	}
	inline Nat shl(const Nat& a, int64_t s)
	{
		if (a.zero() || s == 0) return a;
		size_t w = size_t(s >> 6);
		int b = int(s & 63);
		Nat r;
		r.d.assign(a.d.size() + w + 1, 0);
		for (size_t i = 0; i < a.d.size(); ++i)
		{
			r.d[i + w] |= a.d[i] << b;
			if (b) r.d[i + w + 1] |= a.d[i] >> (64 - b);
		}
		r.trim();
		return r;
	}
	inline Nat shr(const Nat& a, int64_t s)
	{
		if (s <= 0) return s == 0 ? a : shl(a, -s);
		size_t w = size_t(s >> 6);
		int b = int(s & 63);
		Nat r;
		if (w >= a.d.size()) return r;
		r.d.assign(a.d.size() - w, 0);
		for (size_t i = w; i < a.d.size(); ++i)
		{
			u64 v = a.d[i] >> b;
			if (b && i + 1 < a.d.size()) v |= a.d[i + 1] << (64 - b);
			r.d[i - w] = v;
		}
		r.trim();
		return r;
	}
	inline Nat mul(const Nat& a, const Nat& b)
	{
		Nat r;
		if (a.zero() || b.zero()) return r;
		r.d.assign(a.d.size() + b.d.size(), 0);
		for (size_t i = 0; i < a.d.size(); ++i)
		{
			u64 c = 0, x = a.d[i];
			if (!x) continue;
			for (size_t j = 0; j < b.d.size(); ++j)
			{
				u128 t = u128(x) * b.d[j] + r.d[i + j] + c;
				r.d[i + j] = u64(t);
				c = u64(t >> 64);
			}
			r.d[i + b.d.size()] = c;
		}
		r.trim();
		return r;
	}
	inline Nat mul_small(const Nat& a, u64 m)
	{
		Nat r;
		if (a.zero() || !m) return r;
		r.d.resize(a.d.size() + 1);
		u64 c = 0;
		for (size_t i = 0; i < a.d.size(); ++i)
		{
			u128 t = u128(a.d[i]) * m + c;
			r.d[i] = u64(t);
			c = u64(t >> 64);
		}
		r.d[a.d.size()] = c;
		r.trim();
		return r;
	}
	// quotient of a / m, remainder in *rem
	inline Nat div_small(const Nat& a, u64 m, u64* rem = nullptr)
	{
		Nat q;
		q.d.resize(a.d.size());
		u64 r = 0;
		for (size_t i = a.d.size(); i-- > 0;)
		{
			u128 t = (u128(r) << 64) | a.d[i];
			q.d[i] = u64(t / m);
			r = u64(t % m);
		}
		q.trim();
		if (rem) *rem = r;
		return q;
	}
	// Knuth algorithm D: q = floor(a / b), *r = a mod b
	inline Nat divmod(const Nat& a, const Nat& b, Nat* r = nullptr)
	{
		if (b.zero()) throw std::domain_error("qboot_b200: division by zero");
		if (cmp(a, b) < 0)
		{
			if (r) *r = a;
			return Nat();
		}
		if (b.d.size() == 1)
		{
			u64 rem;
			Nat q = div_small(a, b.d[0], &rem);
			if (r) *r = Nat(rem);
			return q;
		}
		const int s = __builtin_clzll(b.d.back());
		Nat u = shl(a, s), v = shl(b, s);
		const size_t n = v.d.size();
		u.d.resize(std::max(u.d.size(), a.d.size() + 1), 0);
		if (u.d.size() < n + 1) u.d.resize(n + 1, 0);
		if (u.d.back() != 0) u.d.push_back(0);
		const size_t m = u.d.size() - n - 1;
		Nat q;
		q.d.assign(m + 1, 0);
		const u64 v1 = v.d[n - 1], v2 = v.d[n - 2];
		for (size_t j = m + 1; j-- > 0;)
		{
			u128 num = (u128(u.d[j + n]) << 64) | u.d[j + n - 1];
			u128 qh = num / v1, rh = num % v1;
			while ((qh >> 64) || (u128(u64(qh)) * v2 > ((rh << 64) | u.d[j + n - 2])))
			{
				--qh;
				rh += v1;
				if (rh >> 64) break;
			}
			// multiply and subtract
			u64 borrow = 0, carry = 0;
			for (size_t i = 0; i < n; ++i)
			{
				u128 p = u128(u64(qh)) * v.d[i] + carry;
				carry = u64(p >> 64);
				u128 t = u128(u.d[i + j]) - u64(p) - borrow;
				u.d[i + j] = u64(t);
				borrow = (t >> 64) ? 1 : 0;
			}
			u128 t = u128(u.d[j + n]) - carry - borrow;
			u.d[j + n] = u64(t);
			if (t >> 64)
			{
				--qh;
				u64 c = 0;
				for (size_t i = 0; i < n; ++i)
				{
					u128 s2 = u128(u.d[i + j]) + v.d[i] + c;
					u.d[i + j] = u64(s2);
					c = u64(s2 >> 64);
				}
				u.d[j + n] += c;
			}
			q.d[j] = u64(qh);
		}
		q.trim();
		if (r)
		{
			u.trim();
			*r = shr(u, s);
		}
		return q;
	}
	inline Nat pow_small(u64 base, uint32_t e)
	{
		Nat r(1), b(base);
		while (e)
		{
			if (e & 1u) r = mul(r, b);
			e >>= 1;
			if (e) b = mul(b, b);
		}
		return r;
	}
	// 10^k, memoised (decimal input / output)
	inline const Nat& pow10(uint32_t k)
	{
		static std::mutex mu;
		static std::map<uint32_t, Nat> memo;
		std::lock_guard<std::mutex> g(mu);
		auto it = memo.find(k);
		if (it == memo.end()) it = memo.emplace(k, pow_small(10, k)).first;
		return it->second;
	}
	// floor(sqrt(a))
	inline Nat isqrt(const Nat& a)
	{
		if (a.zero()) return Nat();
		// Newton from above: x <- (x + a / x) / 2
		Nat x = shl(Nat(1), (a.bits() + 1) / 2);
		for (;;)
		{
			Nat y = shr(add(x, divmod(a, x)), 1);
			if (cmp(y, x) >= 0) return x;
			x = y;
		}
	}

	// ---- working-precision float: value = (-1)^neg * m * 2^e, exact unless truncated -------------------------------------
	struct XF
	{
		bool neg = false;
		Nat m;
		int64_t e = 0;
		bool zero() const { return m.zero(); }
		// exponent of the leading bit: value in [2^(top-1), 2^top)
		int64_t top() const { return m.bits() + e; }
	};
	inline XF xf_from_int(int64_t v)
	{
		XF r;
		r.neg = v < 0;
		r.m = Nat(v < 0 ? u64(0) - u64(v) : u64(v));
		return r;
	}
	// keep the leading wp bits (truncation toward zero: relative error < 2^(1-wp))
	inline void xf_trunc(XF& a, int64_t wp)
	{
		int64_t b = a.m.bits();
		if (b > wp)
		{
			a.m = shr(a.m, b - wp);
			a.e += b - wp;
		}
	}
	inline XF xf_mul(const XF& a, const XF& b, int64_t wp)
	{
		XF r;
		r.m = mul(a.m, b.m);
		r.e = a.e + b.e;
		r.neg = a.neg != b.neg;
		xf_trunc(r, wp);
		return r;
	}
	inline XF xf_mul_small(const XF& a, int64_t k, int64_t wp)
	{
		XF r;
		r.m = mul_small(a.m, k < 0 ? u64(0) - u64(k) : u64(k));
		r.e = a.e;
		r.neg = a.neg != (k < 0);
		xf_trunc(r, wp);
		return r;
	}
	inline XF xf_div_small(const XF& a, u64 k, int64_t wp)
	{
		XF r;
		if (a.zero()) return r;
		int64_t sh = std::max<int64_t>(0, wp + 64 - a.m.bits());
		r.m = div_small(shl(a.m, sh), k);
		r.e = a.e - sh;
		r.neg = a.neg;
		xf_trunc(r, wp);
		return r;
	}
	inline XF xf_div(const XF& a, const XF& b, int64_t wp)
	{
		XF r;
		if (a.zero()) return r;
		int64_t sh = std::max<int64_t>(0, wp + 2 + b.m.bits() - a.m.bits());
		r.m = divmod(shl(a.m, sh), b.m);
		r.e = a.e - sh - b.e;
		r.neg = a.neg != b.neg;
		xf_trunc(r, wp);
		return r;
	}
	inline XF xf_add(const XF& a, const XF& b, int64_t wp)
	{
		if (a.zero()) return b;
		if (b.zero()) return a;
		// align; an operand lying entirely below the other's wp + 2 guard window is truncated first
		XF x = a, y = b;
		if (x.top() < y.top()) std::swap(x, y);
		int64_t lo = x.top() - wp - 3;
		if (y.e < lo)
		{
			int64_t cut = lo - y.e;
			y.m = shr(y.m, cut);
			y.e += cut;
			if (y.zero()) return x;
		}
		if (x.e < lo)
		{
			int64_t cut = lo - x.e;
			x.m = shr(x.m, cut);
			x.e += cut;
		}
		int64_t e = std::min(x.e, y.e);
		Nat xm = shl(x.m, x.e - e), ym = shl(y.m, y.e - e);
		XF r;
		r.e = e;
		if (x.neg == y.neg)
		{
			r.m = add(xm, ym);
			r.neg = x.neg;
		}
		else
		{
			int c = cmp(xm, ym);
			if (c == 0) return XF();
			r.m = c > 0 ? sub(xm, ym) : sub(ym, xm);
			r.neg = c > 0 ? x.neg : y.neg;
		}
		xf_trunc(r, wp);
		return r;
	}
	inline XF xf_neg(XF a)
	{
		if (!a.zero()) a.neg = !a.neg;
		return a;
	}
	inline XF xf_sub(const XF& a, const XF& b, int64_t wp) { return xf_add(a, xf_neg(b), wp); }
	inline double xf_to_double(const XF& a)
	{
		if (a.zero()) return 0.0;
		int64_t b = a.m.bits();
		Nat t = b > 64 ? shr(a.m, b - 64) : a.m;
		double v = std::ldexp(double(t.low64()), int(std::max<int64_t>(-100000, std::min<int64_t>(100000, a.e + (b > 64 ? b - 64 : 0)))));
		return a.neg ? -v : v;
	}
	inline XF xf_from_double(double v)
	{
		XF r;
		if (v == 0.0) return r;
		int ex;
		double f = std::frexp(std::fabs(v), &ex);
		r.m = Nat(u64(std::ldexp(f, 63)));
		r.e = int64_t(ex) - 63;
		r.neg = v < 0;
		return r;
	}

	// ---- constants (fixed point, memoised by precision) -----------------------------------------------------------------
	// floor(2^p * atanh(1/n)) up to a few units, p fraction bits
	inline Nat fx_atanh_inv(u64 n, int64_t p)
	{
		Nat pw = div_small(shl(Nat(1), p), n), sum = pw;
		const u64 n2 = n * n;
		for (u64 k = 1; !pw.zero(); ++k)
		{
			pw = div_small(pw, n2);
			sum = add(sum, div_small(pw, 2 * k + 1));
		}
		return sum;
	}
	inline Nat fx_atan_inv(u64 n, int64_t p)
	{
		Nat pw = div_small(shl(Nat(1), p), n), pos = pw, negs;
		const u64 n2 = n * n;
		for (u64 k = 1; !pw.zero(); ++k)
		{
			pw = div_small(pw, n2);
			Nat t = div_small(pw, 2 * k + 1);
			if (k & 1)
				negs = add(negs, t);
			else
				pos = add(pos, t);
		}
		return sub(pos, negs);
	}
	struct ConstCache
	{
		std::mutex mu;
		std::map<int64_t, XF> memo;
		template <class F>
		XF get(int64_t wp, F&& make)
		{
			// round the request up so that nearby precisions share one evaluation
			int64_t key = (wp + 127) / 128 * 128;
			{
				std::lock_guard<std::mutex> g(mu);
				auto it = memo.find(key);
				if (it != memo.end()) return it->second;
			}
			XF v = make(key);
			std::lock_guard<std::mutex> g(mu);
			return memo.emplace(key, v).first->second;
		}
	};
	// each constant: relative error < 2^-wp
	inline XF const_ln2(int64_t wp)
	{
		static ConstCache cache;
		return cache.get(wp, [](int64_t w) {
			int64_t p = w + 32;
			// ln 2 = 18 atanh(1/26) - 2 atanh(1/4801) + 8 atanh(1/8749)
			Nat s = add(mul_small(fx_atanh_inv(26, p), 18), mul_small(fx_atanh_inv(8749, p), 8));
			s = sub(s, mul_small(fx_atanh_inv(4801, p), 2));
			XF r;
			r.m = s;
			r.e = -p;
			return r;
		});
	}
	inline XF const_pi(int64_t wp)
	{
		static ConstCache cache;
		return cache.get(wp, [](int64_t w) {
			int64_t p = w + 32;
			// pi = 16 atan(1/5) - 4 atan(1/239)
			XF r;
			r.m = sub(mul_small(fx_atan_inv(5, p), 16), mul_small(fx_atan_inv(239, p), 4));
			r.e = -p;
			return r;
		});
	}
	// Euler's constant, Brent-McCurley: gamma = U/V - ln n with A_0 = -ln n, B_0 = 1, B_k = B_{k-1} n^2/k^2,
	// A_k = (A_{k-1} n^2 / k + B_k) / k, U = sum A_k, V = sum B_k; truncation error O(e^(-4n))
	inline XF const_euler(int64_t wp)
	{
		static ConstCache cache;
		return cache.get(wp, [](int64_t w) {
			const int64_t tgt = w + 48;
			// n = 2^q with e^(-4n) < 2^-tgt
			int q = 1;
			while (4.0 * std::ldexp(1.0, q) * 1.4426950408889634 < double(tgt) + 8) ++q;
			const u64 n = u64(1) << q, n2 = n * n;
			const int64_t p = tgt + int64_t(2.0 * double(n) * 1.4426950408889634) + 64;  // terms grow to e^(2n)
			XF l2 = const_ln2(p + 16);
			// A is negative at first: keep sign and magnitude separately
			Nat ln_n = mul_small(shr(l2.m, -(l2.e + p)), u64(q));  // q ln 2 with p fraction bits
			Nat B = shl(Nat(1), p), V = B;
			// signed fixed point via (neg, magnitude)
			bool a_neg = true, u_neg = true;
			Nat A = ln_n, U = ln_n;
			auto add_signed = [](bool& rn, Nat& r, bool xn, const Nat& x) {
				if (rn == xn)
					r = add(r, x);
				else if (cmp(r, x) >= 0)
					r = sub(r, x);
				else
				{
					r = sub(x, r);
					rn = xn;
				}
			};
			for (u64 k = 1;; ++k)
			{
				B = div_small(div_small(mul_small(B, n2), k), k);
				A = div_small(mul_small(A, n2), k);
				add_signed(a_neg, A, false, B);
				A = div_small(A, k);
				add_signed(u_neg, U, a_neg, A);
				V = add(V, B);
				if (A.bits() < 8 && B.bits() < 8) break;
			}
			XF r;
			r.m = divmod(shl(U, tgt + 8), V);
			r.e = -(tgt + 8);
			r.neg = u_neg;
			return r;
		});
	}

	// ---- elementary functions: relative error of the result < 2^(-wp) ------------------------------------------------------
	inline XF xf_exp(const XF& x, int64_t wp)
	{
		if (x.zero()) return xf_from_int(1);
		if (x.top() > 40) throw std::overflow_error("qboot_b200: exp argument out of range");
		// k = round(x / ln 2);  r = x - k ln 2 in [-0.35, 0.35];  exp(r) = (exp(r / 2^s))^(2^s)
		const int64_t s = std::max<int64_t>(4, int64_t(std::sqrt(double(wp)) * 0.7));
		const int64_t w = wp + s + 48 + std::max<int64_t>(0, x.top());
		const double xd = xf_to_double(x);
		const int64_t k = int64_t(std::llround(xd * 1.4426950408889634));
		XF r = x;
		if (k) r = xf_sub(x, xf_mul_small(const_ln2(w + 64), k, w + 64), w);
		if (r.zero())
		{
			XF one = xf_from_int(1);
			one.e += k;
			return one;
		}
		r.e -= s;
		// Taylor in fixed point with w fraction bits (|r| < 2^-s)
		int64_t sh = -(r.e + w);  // r.m * 2^r.e = (r.m >> sh) * 2^-w
		Nat rf = shr(r.m, sh);
		Nat one = shl(Nat(1), w), pos = one, negs, term = one;
		for (u64 n = 1;; ++n)
		{
			term = div_small(shr(mul(term, rf), w), n);
			if (term.zero()) break;
			if (r.neg && (n & 1))
				negs = add(negs, term);
			else
				pos = add(pos, term);
		}
		XF v;
		v.m = sub(pos, negs);
		v.e = -w;
		for (int64_t i = 0; i < s; ++i) v = xf_mul(v, v, w);
		v.e += k;
		xf_trunc(v, wp + 8);
		return v;
	}
	// ln x, x > 0
	inline XF xf_log(const XF& x, int64_t wp)
	{
		if (x.zero() || x.neg) throw std::domain_error("qboot_b200: log of a non-positive number");
		// x = m 2^t with m in [1/sqrt2, sqrt2):  ln x = t ln 2 + ln m, so the only cancellation left is x ~ 1 (t = 0), where
		// ln m itself is small and is computed to relative accuracy.
		// ln m by Newton on exp:  y += 2 (m - e^y) / (m + e^y)  (cubic: absolute error 2^-a -> about 2^-3a)
		int64_t t = x.top();
		XF m = x;
		m.e -= t;
		if (xf_to_double(m) < 0.7071067811865476)
		{
			m.e += 1;
			t -= 1;
		}
		const XF u = xf_sub(m, xf_from_int(1), int64_t(1) << 40);  // exact
		if (u.zero())
		{
			if (t == 0) return XF();
			return xf_mul_small(const_ln2(wp + 16), t, wp + 8);
		}
		const int64_t k = std::max<int64_t>(0, -u.top());  // |ln m| ~ 2^-k
		XF y;
		int64_t a;
		if (k >= 20)
		{
			y = u;  // ln(1 + u) = u - u^2/2 + ...
			xf_trunc(y, 2 * k + 64);
			a = 2 * k - 2;
		}
		else
		{
			y = xf_from_double(std::log(xf_to_double(m)));
			a = 44;
		}
		const int64_t target = wp + 24 + (t == 0 ? k + 2 : 0);
		while (a < target)
		{
			const int64_t cur = std::min(target, 3 * a) + 40;
			XF ey = xf_exp(y, cur);
			XF num = xf_sub(m, ey, cur), den = xf_add(m, ey, cur);
			XF dlt = xf_div(num, den, cur);
			dlt.e += 1;
			y = xf_add(y, dlt, cur);
			a = 3 * a - 6;
		}
		if (t == 0) return y;
		return xf_add(y, xf_mul_small(const_ln2(wp + 72), t, wp + 72), wp + 16);
	}
	// E1(x) = Gamma(0, x), x > 0:  -gamma - ln x + sum_{k >= 1} (-1)^(k+1) x^k / (k k!)
	inline XF xf_e1(const XF& x, int64_t wp)
	{
		if (x.zero() || x.neg) throw std::domain_error("qboot_b200: gamma_inc(0, x) needs x > 0");
		const double xd = xf_to_double(x);
		if (xd > 4000.0) throw std::overflow_error("qboot_b200: gamma_inc(0, x) argument too large");
		// terms reach e^x / x while the result is ~ e^-x / x: carry 2 x log2(e) extra bits
		const int64_t extra = int64_t(2.0 * xd * 1.4426950408889634) + 16;
		const int64_t w = wp + extra + 64;
		// fixed point with w fraction bits; integer part of terms up to e^x
		int64_t sh = -(x.e + w);
		Nat xf = sh >= 0 ? shr(x.m, sh) : shl(x.m, -sh);
		Nat term = shl(Nat(1), w), pos, negs;  // term = x^k / k!
		for (u64 k = 1;; ++k)
		{
			term = div_small(shr(mul(term, xf), w), k);
			if (term.zero()) break;
			Nat t = div_small(term, k);
			if (k & 1)
				pos = add(pos, t);
			else
				negs = add(negs, t);
		}
		XF s;
		if (cmp(pos, negs) >= 0)
			s.m = sub(pos, negs);
		else
		{
			s.m = sub(negs, pos);
			s.neg = true;
		}
		s.e = -w;
		XF r = xf_sub(s, const_euler(w), w + 8);
		r = xf_sub(r, xf_log(x, w), w + 8);
		return r;
	}
}  // namespace qb::host
