// qboot_b200 host library -- arithmetic on number records for the C++ façade (declared in include/qboot/b200/hostops.hpp).
//
// Basic operations dispatch on the limb count to the engine's own `__host__ __device__` routines (csrc/mpf.cuh,
// csrc/hostmath.hpp) compiled for the host.  Operations with rational operands, transcendentals and decimal conversion
// go through exact integer arithmetic (bignum.hpp) and are rounded once; every result is the correctly rounded one, the
// contract of the MPFR calls they replace (include/qboot/mp/real.hpp:243-255,324-361,688-804,987-1030).
#include "qboot/b200/hostops.hpp"

#include <cstring>
#include <stdexcept>

#include "../csrc/hostmath.hpp"
#include "bignum.hpp"

namespace qb::host
{
	namespace
	{
#define QB_HOST_LIMBS(X) X(16) X(19) X(25) X(32) X(38) X(48)
#define QB_DISPATCH(LV, ...)                                        \
	switch (LV)                                                     \
	{                                                               \
	case 16: { constexpr int L = 16; __VA_ARGS__; } break;          \
	case 19: { constexpr int L = 19; __VA_ARGS__; } break;          \
	case 25: { constexpr int L = 25; __VA_ARGS__; } break;          \
	case 32: { constexpr int L = 32; __VA_ARGS__; } break;          \
	case 38: { constexpr int L = 38; __VA_ARGS__; } break;          \
	case 48: { constexpr int L = 48; __VA_ARGS__; } break;          \
	default: throw std::invalid_argument("qboot_b200: unsupported precision"); \
	}
		template <int L>
		const mpf<L>& in(const uint32_t* p)
		{
			return *reinterpret_cast<const mpf<L>*>(p);
		}
		template <int L>
		void out(uint32_t* p, const mpf<L>& v)
		{
			std::memcpy(p, &v, sizeof(v));
		}
		inline int LW(int prec) { return (prec + 31) / 32; }
		inline uint32_t kind_of(const uint32_t* a) { return a[0] & 3u; }
		inline bool neg_of(const uint32_t* a) { return (a[0] >> 31) != 0; }
		void set_special(uint32_t* r, int prec, uint32_t kind, bool neg)
		{
			std::memset(r, 0, sizeof(uint32_t) * size_t(LW(prec) + 2));
			r[0] = kind | (neg ? SIGN_BIT : 0u);
		}

		// record -> exact (m, e)
		XF to_xf(const uint32_t* a, int prec)
		{
			XF v;
			if (kind_of(a) != K_REG) return v;
			const int L = LW(prec);
			v.m.d.assign(size_t(L + 1) / 2, 0);
			for (int i = 0; i < L; ++i) v.m.d[size_t(i / 2)] |= u64(a[2 + i]) << (32 * (i & 1));
			v.m.trim();
			v.e = int64_t(int32_t(a[1])) - 32 * int64_t(L);
			v.neg = neg_of(a);
			return v;
		}
		// RN(|v| + (sticky ? epsilon : 0)) with sign -> record
		void from_xf(uint32_t* r, const XF& v, bool sticky, int prec)
		{
			const int L = LW(prec);
			if (v.zero())
			{
				set_special(r, prec, K_ZERO, false);
				return;
			}
			Nat q = v.m;
			int64_t e = v.e;
			int64_t b = q.bits();
			if (b > prec)
			{
				const int64_t drop = b - prec;
				const bool rnd = q.bit(drop - 1);
				const bool st = sticky || q.any_below(drop - 1);
				q = shr(q, drop);
				e += drop;
				if (rnd && (st || (q.low64() & 1u)))
				{
					q = add(q, Nat(1));
					if (q.bits() > prec)
					{
						q = shr(q, 1);
						e += 1;
					}
				}
				b = prec;
			}
			// left-align in 32 L bits
			Nat al = shl(q, 32 * int64_t(L) - b);
			const int64_t ex = e + b;
			if (ex > INT32_MAX || ex < INT32_MIN) throw std::overflow_error("qboot_b200: exponent out of range");
			std::memset(r, 0, sizeof(uint32_t) * size_t(L + 2));
			r[0] = K_REG | (v.neg ? SIGN_BIT : 0u);
			r[1] = uint32_t(int32_t(ex));
			for (int i = 0; i < L; ++i)
			{
				size_t w = size_t(i / 2);
				u64 limb = w < al.d.size() ? al.d[w] : 0;
				r[2 + i] = uint32_t(limb >> (32 * (i & 1)));
			}
		}
		// RN(num / den * 2^e2) with sign
		void round_ratio(uint32_t* r, bool neg, const Nat& num, const Nat& den, int64_t e2, int prec)
		{
			if (num.zero())
			{
				set_special(r, prec, K_ZERO, false);
				return;
			}
			const int64_t s = std::max<int64_t>(0, int64_t(prec) + 2 + den.bits() - num.bits());
			Nat rem;
			XF v;
			v.m = divmod(shl(num, s), den, &rem);
			v.e = e2 - s;
			v.neg = neg;
			from_xf(r, v, !rem.zero(), prec);
		}
		// v approximates the true value with relative error < 2^-acc: round to nearest if that is provably the correct
		// rounding (the true value cannot lie on the other side of a midpoint)
		bool ziv_round(uint32_t* r, const XF& v, int64_t acc, int prec)
		{
			if (v.zero()) return false;
			const int64_t b = v.m.bits(), drop = b - prec;
			const int64_t errbits = b - acc + 1;  // error < 2^errbits units of v's last bit
			if (drop < errbits + 2) return false;
			// low = v mod 2^drop; ambiguous iff |low - 2^(drop-1)| < 2^errbits
			Nat low = sub(v.m, shl(shr(v.m, drop), drop)), half = shl(Nat(1), drop - 1);
			Nat dist = cmp(low, half) >= 0 ? sub(low, half) : sub(half, low);
			if (dist.bits() <= errbits + 1) return false;
			from_xf(r, v, true, prec);
			return true;
		}
		// run f(wp) -> XF with relative error < 2^-(wp - 8) until the rounding is certain
		template <class F>
		void ziv(uint32_t* r, int prec, F&& f)
		{
			int64_t wp = prec + 64;
			XF v;
			for (int tries = 0; tries < 7; ++tries, wp += int64_t(64) << tries)
			{
				v = f(wp);
				if (v.zero())
				{
					set_special(r, prec, K_ZERO, false);
					return;
				}
				if (ziv_round(r, v, wp - 8, prec)) return;
			}
			// indistinguishable from a midpoint at ~8000 extra bits: the value IS that midpoint; ties go to even
			const int64_t drop = v.m.bits() - prec;
			XF m;
			m.m = add(shr(v.m, drop - 1), Nat(shr(v.m, drop - 2).low64() & 1u));  // nearest half-ulp multiple
			m.e = v.e + drop - 1;
			m.neg = v.neg;
			from_xf(r, m, false, prec);
		}
		XF xf_of_si(int64_t v) { return xf_from_int(v); }
	}  // namespace

	int limbs_for(int prec)
	{
		if (prec < 2) return 0;
		const int L = LW(prec);
		switch (L)
		{
#define X(n) case n:
			QB_HOST_LIMBS(X)
#undef X
			return L;
		default: return 0;
		}
	}
	void check_prec(int prec)
	{
		if (!limbs_for(prec))
			throw std::invalid_argument("qboot_b200: mp::global_prec = " + std::to_string(prec) +
			                            " has no compiled kernels (supported limb counts: 16, 19, 25, 32, 38, 48 x 32 bits)");
	}

	// ---- basic operations -------------------------------------------------------------------------------------------
	void h_add(uint32_t* r, const uint32_t* a, const uint32_t* b, int prec)
	{
		QB_DISPATCH(LW(prec), mpf<L> t; add(t, in<L>(a), in<L>(b), prec); out(r, t))
	}
	void h_sub(uint32_t* r, const uint32_t* a, const uint32_t* b, int prec)
	{
		QB_DISPATCH(LW(prec), mpf<L> t; add(t, in<L>(a), in<L>(b), prec, true); out(r, t))
	}
	void h_mul(uint32_t* r, const uint32_t* a, const uint32_t* b, int prec)
	{
		QB_DISPATCH(LW(prec), mpf<L> t; mul(t, in<L>(a), in<L>(b), prec); out(r, t))
	}
	void h_div(uint32_t* r, const uint32_t* a, const uint32_t* b, int prec)
	{
		QB_DISPATCH(LW(prec), mpf<L> t; div(t, in<L>(a), in<L>(b), prec); out(r, t))
	}
	void h_fma(uint32_t* r, const uint32_t* a, const uint32_t* b, const uint32_t* c, int prec)
	{
		QB_DISPATCH(LW(prec), mpf<L> t; fma(t, in<L>(a), in<L>(b), in<L>(c), prec); out(r, t))
	}
	void h_fms(uint32_t* r, const uint32_t* a, const uint32_t* b, const uint32_t* c, int prec)
	{
		QB_DISPATCH(LW(prec), mpf<L> t, nc = in<L>(c); neg(nc); fma(t, in<L>(a), in<L>(b), nc, prec); out(r, t))
	}
	void h_sqrt(uint32_t* r, const uint32_t* a, int prec)
	{
		QB_DISPATCH(LW(prec), mpf<L> t; sqrt(t, in<L>(a), prec); out(r, t))
	}
	void h_neg(uint32_t* r, const uint32_t* a, int prec)
	{
		const size_t W = size_t(LW(prec) + 2);
		if (r != a) std::memmove(r, a, 4 * W);
		if (kind_of(r) != K_NAN) r[0] ^= SIGN_BIT;
	}
	void h_abs(uint32_t* r, const uint32_t* a, int prec)
	{
		const size_t W = size_t(LW(prec) + 2);
		if (r != a) std::memmove(r, a, 4 * W);
		r[0] &= ~SIGN_BIT;
	}
	void h_set_si(uint32_t* r, int64_t v, int prec)
	{
		QB_DISPATCH(LW(prec), mpf<L> t; set_si(t, v, prec); out(r, t))
	}
	void h_set_ui(uint32_t* r, uint64_t v, int prec)
	{
		QB_DISPATCH(LW(prec), mpf<L> t; set_ui_2exp(t, v, 0, prec); out(r, t))
	}
	void h_set_si_2exp(uint32_t* r, int64_t v, int64_t e, int prec)
	{
		XF x = xf_of_si(v);
		x.e = e;
		from_xf(r, x, false, prec);
	}
	void h_set_d(uint32_t* r, double v, int prec)
	{
		if (v != v) return set_special(r, prec, K_NAN, false);
		if (std::isinf(v)) return set_special(r, prec, K_INF, v < 0);
		if (v == 0.0) return set_special(r, prec, K_ZERO, std::signbit(v));
		from_xf(r, xf_from_double(v), false, prec);
	}
	void h_set_q(uint32_t* r, int64_t num, int64_t den, int prec)
	{
		if (den < 0)
		{
			num = -num;
			den = -den;
		}
		round_ratio(r, num < 0, Nat(num < 0 ? u64(0) - u64(num) : u64(num)), Nat(u64(den)), 0, prec);
	}
	void h_mul_si(uint32_t* r, const uint32_t* a, int64_t v, int prec)
	{
		QB_DISPATCH(LW(prec), mpf<L> t; mul_si(t, in<L>(a), v, prec); out(r, t))
	}
	void h_div_si(uint32_t* r, const uint32_t* a, int64_t v, int prec)
	{
		if (v > INT32_MAX || v < -INT32_MAX)
		{
			uint32_t t[kMaxWords];
			h_set_si(t, v, prec);
			return h_div(r, a, t, prec);
		}
		QB_DISPATCH(LW(prec), mpf<L> t; div_si(t, in<L>(a), v, prec); out(r, t))
	}
	void h_add_si(uint32_t* r, const uint32_t* a, int64_t v, int prec)
	{
		QB_DISPATCH(LW(prec), mpf<L> t; add_si(t, in<L>(a), v, prec); out(r, t))
	}
	void h_si_sub(uint32_t* r, int64_t v, const uint32_t* a, int prec)
	{
		QB_DISPATCH(LW(prec), mpf<L> t; si_sub(t, v, in<L>(a), prec); out(r, t))
	}
	void h_si_div(uint32_t* r, int64_t v, const uint32_t* a, int prec)
	{
		uint32_t t[kMaxWords];
		h_set_si(t, v, prec);
		h_div(r, t, a, prec);
	}
	void h_mul_2si(uint32_t* r, const uint32_t* a, int64_t e, int prec)
	{
		const size_t W = size_t(LW(prec) + 2);
		if (r != a) std::memmove(r, a, 4 * W);
		if (kind_of(r) == K_REG) r[1] = uint32_t(int32_t(int64_t(int32_t(r[1])) + e));
	}
	// rational operands: exact integer arithmetic, one rounding (mpfr_mul_q / mpfr_div_q / mpfr_add_q / mpfr_sub_q)
	namespace
	{
		void split_q(int64_t& num, int64_t& den, bool& neg, Nat& n, Nat& d)
		{
			if (den == 0) throw std::domain_error("qboot_b200: rational with zero denominator");
			if (den < 0)
			{
				num = -num;
				den = -den;
			}
			neg = num < 0;
			n = Nat(neg ? u64(0) - u64(num) : u64(num));
			d = Nat(u64(den));
		}
	}  // namespace
	void h_mul_q(uint32_t* r, const uint32_t* a, int64_t num, int64_t den, int prec)
	{
		bool qn;
		Nat n, d;
		split_q(num, den, qn, n, d);
		if (kind_of(a) != K_REG)
		{
			// 0 * q = +-0, inf * q = +-inf (q != 0), nan
			if (kind_of(a) == K_NAN || (kind_of(a) == K_INF && n.zero())) return set_special(r, prec, K_NAN, false);
			return set_special(r, prec, kind_of(a), neg_of(a) != qn);
		}
		if (n.zero()) return set_special(r, prec, K_ZERO, neg_of(a) != qn);
		XF x = to_xf(a, prec);
		round_ratio(r, x.neg != qn, mul(x.m, n), d, x.e, prec);
	}
	void h_div_q(uint32_t* r, const uint32_t* a, int64_t num, int64_t den, int prec)
	{
		bool qn;
		Nat n, d;
		split_q(num, den, qn, n, d);
		if (kind_of(a) != K_REG)
		{
			if (kind_of(a) == K_NAN || (kind_of(a) == K_ZERO && n.zero())) return set_special(r, prec, K_NAN, false);
			return set_special(r, prec, kind_of(a), neg_of(a) != qn);
		}
		if (n.zero()) return set_special(r, prec, K_INF, neg_of(a) != qn);
		XF x = to_xf(a, prec);
		round_ratio(r, x.neg != qn, mul(x.m, d), n, x.e, prec);
	}
	void h_add_q(uint32_t* r, const uint32_t* a, int64_t num, int64_t den, int prec)
	{
		bool qn;
		Nat n, d;
		split_q(num, den, qn, n, d);
		if (kind_of(a) == K_NAN || kind_of(a) == K_INF)
		{
			const size_t W = size_t(LW(prec) + 2);
			if (r != a) std::memmove(r, a, 4 * W);
			return;
		}
		if (n.zero())
		{
			// mpfr_add_q with q = 0 returns a itself (the sign of a zero is kept)
			const size_t W = size_t(LW(prec) + 2);
			if (r != a) std::memmove(r, a, 4 * W);
			return;
		}
		if (kind_of(a) == K_ZERO) return round_ratio(r, qn, n, d, 0, prec);
		// a + n/d = (m 2^e d + n) / d; bring both to the common power of two
		XF x = to_xf(a, prec);
		Nat am = mul(x.m, d), qm = n;
		int64_t e = x.e;
		if (e >= 0)
		{
			am = shl(am, e);
			e = 0;
		}
		else
			qm = shl(qm, -e);
		bool rn;
		Nat s;
		if (x.neg == qn)
		{
			s = add(am, qm);
			rn = x.neg;
		}
		else
		{
			int c = cmp(am, qm);
			if (c == 0) return set_special(r, prec, K_ZERO, false);
			s = c > 0 ? sub(am, qm) : sub(qm, am);
			rn = c > 0 ? x.neg : qn;
		}
		round_ratio(r, rn, s, d, e, prec);
	}
	void h_sub_q(uint32_t* r, const uint32_t* a, int64_t num, int64_t den, int prec)
	{
		if (num == INT64_MIN) throw std::overflow_error("qboot_b200: rational out of range");
		h_add_q(r, a, -num, den, prec);
	}
	int h_cmp(const uint32_t* a, const uint32_t* b, int prec)
	{
		const uint32_t ka = kind_of(a), kb = kind_of(b);
		if (ka == K_NAN || kb == K_NAN) return 0;
		if (ka == K_INF || kb == K_INF)
		{
			int va = ka == K_INF ? (neg_of(a) ? -2 : 2) : 0, vb = kb == K_INF ? (neg_of(b) ? -2 : 2) : 0;
			if (ka != K_INF) va = kind_of(a) == K_ZERO ? 0 : (neg_of(a) ? -1 : 1);
			if (kb != K_INF) vb = kind_of(b) == K_ZERO ? 0 : (neg_of(b) ? -1 : 1);
			return va < vb ? -1 : va > vb ? 1 : 0;
		}
		int res = 0;
		QB_DISPATCH(LW(prec), res = cmp(in<L>(a), in<L>(b)))
		return res;
	}
	int h_cmpabs(const uint32_t* a, const uint32_t* b, int prec)
	{
		uint32_t x[kMaxWords], y[kMaxWords];
		h_abs(x, a, prec);
		h_abs(y, b, prec);
		return h_cmp(x, y, prec);
	}
	int h_cmp_si(const uint32_t* a, int64_t v, int prec)
	{
		uint32_t t[kMaxWords];
		h_set_si(t, v, prec);
		return h_cmp(a, t, prec);
	}
	int h_cmp_q(const uint32_t* a, int64_t num, int64_t den, int prec)
	{
		// sign of a - num/den, exactly: compare a * den with num
		if (den < 0)
		{
			num = -num;
			den = -den;
		}
		if (kind_of(a) == K_NAN) return 0;
		if (kind_of(a) == K_INF) return neg_of(a) ? -1 : 1;
		const int sa = kind_of(a) == K_ZERO ? 0 : (neg_of(a) ? -1 : 1), sq = num < 0 ? -1 : num > 0 ? 1 : 0;
		if (sa != sq) return sa < sq ? -1 : 1;
		if (sa == 0) return 0;
		XF x = to_xf(a, prec);
		Nat l = mul(x.m, Nat(u64(den))), rr = Nat(num < 0 ? u64(0) - u64(num) : u64(num));
		if (x.e >= 0)
			l = shl(l, x.e);
		else
			rr = shl(rr, -x.e);
		return sa * cmp(l, rr);
	}
	bool h_is_integer(const uint32_t* a, int prec)
	{
		bool res = false;
		QB_DISPATCH(LW(prec), res = is_integer(in<L>(a)))
		return res;
	}
	namespace
	{
		// integer rounding: mode 0 floor, 1 ceil, 2 round-half-away (mpfr_round)
		void to_integer(uint32_t* r, const uint32_t* a, int prec, int mode)
		{
			if (kind_of(a) != K_REG)
			{
				const size_t W = size_t(LW(prec) + 2);
				if (r != a) std::memmove(r, a, 4 * W);
				return;
			}
			XF x = to_xf(a, prec);
			if (x.e >= 0)
			{
				const size_t W = size_t(LW(prec) + 2);
				if (r != a) std::memmove(r, a, 4 * W);
				return;
			}
			const int64_t fr = -x.e;
			Nat ip = shr(x.m, fr);
			const bool frac = x.m.any_below(fr), halfbit = x.m.bit(fr - 1);
			bool up = false;
			if (mode == 0) up = x.neg && frac;
			if (mode == 1) up = !x.neg && frac;
			if (mode == 2) up = halfbit;
			if (up) ip = add(ip, Nat(1));
			if (ip.zero()) return set_special(r, prec, K_ZERO, x.neg);
			XF v;
			v.m = ip;
			v.neg = x.neg;
			from_xf(r, v, false, prec);
		}
	}  // namespace
	void h_floor(uint32_t* r, const uint32_t* a, int prec) { to_integer(r, a, prec, 0); }
	void h_ceil(uint32_t* r, const uint32_t* a, int prec) { to_integer(r, a, prec, 1); }
	void h_round(uint32_t* r, const uint32_t* a, int prec) { to_integer(r, a, prec, 2); }
	double h_get_d(const uint32_t* a, int prec)
	{
		if (kind_of(a) == K_NAN) return std::nan("");
		if (kind_of(a) == K_INF) return neg_of(a) ? -HUGE_VAL : HUGE_VAL;
		return xf_to_double(to_xf(a, prec));
	}
	int64_t h_get_si(const uint32_t* a, int prec)
	{
		if (kind_of(a) != K_REG) return 0;
		XF x = to_xf(a, prec);
		Nat ip = x.e >= 0 ? shl(x.m, x.e) : shr(x.m, -x.e);
		if (ip.bits() > 63) throw std::overflow_error("qboot_b200: integer conversion out of range");
		int64_t v = int64_t(ip.low64());
		return x.neg ? -v : v;
	}

	// ---- correctly rounded transcendentals -----------------------------------------------------------------------------
	void h_const_pi(uint32_t* r, int prec)
	{
		ziv(r, prec, [](int64_t wp) { return const_pi(wp); });
	}
	void h_const_log2(uint32_t* r, int prec)
	{
		ziv(r, prec, [](int64_t wp) { return const_ln2(wp); });
	}
	void h_exp(uint32_t* r, const uint32_t* a, int prec)
	{
		switch (kind_of(a))
		{
		case K_NAN: return set_special(r, prec, K_NAN, false);
		case K_INF: return set_special(r, prec, neg_of(a) ? K_ZERO : K_INF, false);
		case K_ZERO: return h_set_si(r, 1, prec);
		default: break;
		}
		const XF x = to_xf(a, prec);
		if (x.top() < -int64_t(prec) - 2)
		{
			// exp(x) = 1 + x + ...: rounds to 1 or its neighbour
			XF v = xf_from_int(1);
			v = xf_add(v, x, 4 * int64_t(prec) + 64 - x.top());
			return from_xf(r, v, false, prec);
		}
		ziv(r, prec, [&](int64_t wp) { return xf_exp(x, wp); });
	}
	void h_log(uint32_t* r, const uint32_t* a, int prec)
	{
		switch (kind_of(a))
		{
		case K_NAN: return set_special(r, prec, K_NAN, false);
		case K_INF: return set_special(r, prec, neg_of(a) ? K_NAN : K_INF, false);
		case K_ZERO: return set_special(r, prec, K_INF, true);
		default: break;
		}
		if (neg_of(a)) return set_special(r, prec, K_NAN, false);
		const XF x = to_xf(a, prec);
		ziv(r, prec, [&](int64_t wp) { return xf_log(x, wp); });
	}
	namespace
	{
		// ln of the last base, per thread: a scale factor raises the same base (4 rho) to some eighty powers, and the
		// logarithm is two thirds of a power's cost.  A cached value of higher working precision serves any request of lower
		// precision -- it only satisfies the error bound better, and the Ziv loop around it makes the rounded result unique.
		XF cached_log(const XF& ax, int64_t wp)
		{
			struct Cache
			{
				Nat m;
				int64_t e = 0, wp = -1;
				XF value;
			};
			thread_local Cache c;
			if (c.wp >= wp && c.e == ax.e && cmp(c.m, ax.m) == 0) return c.value;
			c.value = xf_log(ax, wp + 64);
			c.m = ax.m;
			c.e = ax.e;
			c.wp = wp + 64;
			return c.value;
		}
		// |x|^y for regular x != 1 and regular y through exp(y ln|x|)
		void pow_general(uint32_t* r, const XF& x, const XF& y, bool neg_result, int prec)
		{
			XF ax = x;
			ax.neg = false;
			ziv(r, prec, [&](int64_t wp) {
				XF l0 = cached_log(ax, 64);
				if (l0.zero()) return xf_from_int(1);
				const int64_t tb = std::max<int64_t>(0, l0.top() + y.top()) + 8;  // bits of |y ln x|
				XF lx = cached_log(ax, wp + tb + 8);
				XF t = xf_mul(y, lx, wp + tb + 8);
				XF v = xf_exp(t, wp);
				v.neg = neg_result;
				return v;
			});
		}
	}  // namespace
	void h_pow_si(uint32_t* r, const uint32_t* a, int64_t n, int prec)
	{
		if (n == 0) return h_set_si(r, 1, prec);
		const uint32_t k = kind_of(a);
		const bool odd = (n & 1) != 0, rn = neg_of(a) && odd;
		if (k == K_NAN) return set_special(r, prec, K_NAN, false);
		if (k == K_INF) return set_special(r, prec, n > 0 ? K_INF : K_ZERO, rn);
		if (k == K_ZERO) return set_special(r, prec, n > 0 ? K_ZERO : K_INF, rn);
		XF x = to_xf(a, prec);
		// strip trailing zeros: x = odd_part * 2^e
		int64_t tz = 0;
		while (!x.m.bit(tz)) ++tz;
		x.m = shr(x.m, tz);
		x.e += tz;
		const u64 un = n < 0 ? u64(0) - u64(n) : u64(n);
		const int64_t ob = x.m.bits();
		if (ob == 1)
		{
			// a power of two: exact
			XF v = xf_from_int(1);
			v.e = x.e * n;
			v.neg = rn;
			return from_xf(r, v, false, prec);
		}
		if (un < (u64(1) << 20) && int64_t(un) * ob <= 4 * int64_t(prec) + 64)
		{
			// exact integer power, rounded once (this range contains every exactly representable result and every
			// possible rounding tie: those need odd_part^|n| to fit prec + 1 bits)
			Nat p = Nat(1), b = x.m;
			for (u64 e = un; e; e >>= 1)
			{
				if (e & 1u) p = mul(p, b);
				if (e > 1) b = mul(b, b);
			}
			if (n > 0)
			{
				XF v;
				v.m = p;
				v.e = x.e * n;
				v.neg = rn;
				return from_xf(r, v, false, prec);
			}
			return round_ratio(r, rn, Nat(1), p, x.e * n, prec);
		}
		XF y = xf_from_int(n);
		pow_general(r, x, y, rn, prec);
	}
	void h_pow(uint32_t* r, const uint32_t* a, const uint32_t* b, int prec)
	{
		const uint32_t ka = kind_of(a), kb = kind_of(b);
		if (kb == K_ZERO) return h_set_si(r, 1, prec);
		if (ka == K_NAN || kb == K_NAN) return set_special(r, prec, K_NAN, false);
		if (kb == K_REG && h_is_integer(b, prec))
		{
			XF y = to_xf(b, prec);
			if (y.top() <= 40) return h_pow_si(r, a, h_get_si(b, prec), prec);
		}
		if (ka == K_REG && !neg_of(a) && h_cmp_si(a, 1, prec) == 0) return h_set_si(r, 1, prec);
		if (kb == K_INF)
		{
			// |a| < 1: +inf -> 0, -inf -> inf; |a| > 1 the other way round
			uint32_t one[kMaxWords];
			h_set_si(one, 1, prec);
			const int c = ka == K_ZERO ? -1 : ka == K_INF ? 1 : h_cmpabs(a, one, prec);
			if (c == 0) return h_set_si(r, 1, prec);
			return set_special(r, prec, ((c > 0) != neg_of(b)) ? K_INF : K_ZERO, false);
		}
		if (ka == K_ZERO) return set_special(r, prec, neg_of(b) ? K_INF : K_ZERO, false);
		if (ka == K_INF) return set_special(r, prec, neg_of(b) ? K_ZERO : K_INF, false);
		if (neg_of(a)) return set_special(r, prec, K_NAN, false);  // negative base, non-integer exponent
		pow_general(r, to_xf(a, prec), to_xf(b, prec), false, prec);
	}
	void h_ui_pow(uint32_t* r, uint64_t base, const uint32_t* b, int prec)
	{
		uint32_t t[kMaxWords];
		h_set_ui(t, base, prec);
		h_pow(r, t, b, prec);
	}
	void h_gamma_inc(uint32_t* r, const uint32_t* a, const uint32_t* x, int prec)
	{
		if (kind_of(a) != K_ZERO) throw std::domain_error("qboot_b200: gamma_inc(a, x) is implemented for a = 0 only");
		switch (kind_of(x))
		{
		case K_NAN: return set_special(r, prec, K_NAN, false);
		case K_INF: return set_special(r, prec, neg_of(x) ? K_NAN : K_ZERO, false);
		case K_ZERO: return set_special(r, prec, K_INF, false);
		default: break;
		}
		if (neg_of(x)) return set_special(r, prec, K_NAN, false);
		const XF xv = to_xf(x, prec);
		ziv(r, prec, [&](int64_t wp) { return xf_e1(xv, wp); });
	}

	// ---- decimal text ------------------------------------------------------------------------------------------------
	bool h_from_string(uint32_t* r, std::string_view s, int prec)
	{
		size_t i = 0;
		while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n')) ++i;
		bool neg = false;
		if (i < s.size() && (s[i] == '+' || s[i] == '-')) neg = s[i++] == '-';
		auto ieq = [&](std::string_view w) {
			if (s.size() - i != w.size()) return false;
			for (size_t k = 0; k < w.size(); ++k)
				if (std::tolower(static_cast<unsigned char>(s[i + k])) != w[k]) return false;
			return true;
		};
		if (ieq("inf") || ieq("infinity") || ieq("@inf@"))
		{
			set_special(r, prec, K_INF, neg);
			return true;
		}
		if (ieq("nan") || ieq("@nan@"))
		{
			set_special(r, prec, K_NAN, false);
			return true;
		}
		Nat mant;
		int64_t scale = 0;
		bool digits = false, point = false;
		u64 chunk = 0, chunk_pow = 1;
		auto flush = [&]() {
			if (chunk_pow > 1)
			{
				mant = add(mul_small(mant, chunk_pow), Nat(chunk));
				chunk = 0;
				chunk_pow = 1;
			}
		};
		for (; i < s.size() && s[i] != 'e' && s[i] != 'E' && s[i] != '@'; ++i)
		{
			if (s[i] == '.')
			{
				if (point) return false;
				point = true;
				continue;
			}
			if (s[i] < '0' || s[i] > '9') return false;
			chunk = chunk * 10 + u64(s[i] - '0');
			chunk_pow *= 10;
			if (chunk_pow == 1000000000000000000ull) flush();
			digits = true;
			if (point) --scale;
		}
		flush();
		if (!digits) return false;
		if (i < s.size())
		{
			++i;
			bool eneg = false;
			if (i < s.size() && (s[i] == '+' || s[i] == '-')) eneg = s[i++] == '-';
			if (i >= s.size()) return false;
			int64_t ev = 0;
			for (; i < s.size(); ++i)
			{
				if (s[i] < '0' || s[i] > '9') return false;
				ev = ev * 10 + (s[i] - '0');
				if (ev > 100000000) return false;
			}
			scale += eneg ? -ev : ev;
		}
		if (mant.zero())
		{
			set_special(r, prec, K_ZERO, neg);
			return true;
		}
		if (scale >= 0)
			round_ratio(r, neg, mul(mant, pow10(uint32_t(scale))), Nat(1), 0, prec);
		else
			round_ratio(r, neg, mant, pow10(uint32_t(-scale)), 0, prec);
		return true;
	}
	std::string h_get_digits(const uint32_t* a, int prec, size_t ndigits, int64_t* e10)
	{
		if (kind_of(a) != K_REG || ndigits == 0) throw std::invalid_argument("qboot_b200: h_get_digits needs a regular number");
		const XF x = to_xf(a, prec);
		const int64_t n = int64_t(ndigits);
		// first guess of the decimal exponent d (|a| = 0.ddd * 10^d): floor(log10|a|) + 1
		int64_t d = int64_t(std::floor(double(x.top() - 1) * 0.30102999566398120)) + 1;
		const Nat& lim_hi = pow10(uint32_t(n));
		for (int attempt = 0; attempt < 4; ++attempt)
		{
			// N = RN(|a| * 10^(n - d)) = RN(A / B)
			const int64_t k = n - d;
			Nat A = x.m, B = Nat(1);
			if (k >= 0)
				A = mul(A, pow10(uint32_t(k)));
			else
				B = pow10(uint32_t(-k));
			if (x.e >= 0)
				A = shl(A, x.e);
			else
				B = shl(B, -x.e);
			Nat rem, N = divmod(A, B, &rem);
			// the unrounded quotient decides the exponent: 10^(n-1) <= A/B < 10^n
			if (cmp(N, lim_hi) >= 0)
			{
				++d;
				continue;
			}
			if (cmp(N, pow10(uint32_t(n - 1))) < 0)
			{
				--d;
				continue;
			}
			const int c = cmp(shl(rem, 1), B);
			if (c > 0 || (c == 0 && (N.low64() & 1u))) N = add(N, Nat(1));
			if (cmp(N, lim_hi) >= 0)
			{
				// 99...9.6 rounded up to 10^n
				N = pow10(uint32_t(n - 1));
				++d;
			}
			std::string out(ndigits, '0');
			size_t pos = ndigits;
			while (!N.zero())
			{
				u64 rem19;
				N = div_small(N, 10000000000000000000ull, &rem19);
				for (int j = 0; j < 19 && pos > 0; ++j)
				{
					out[--pos] = char('0' + rem19 % 10);
					rem19 /= 10;
					if (N.zero() && rem19 == 0) break;
				}
			}
			*e10 = d;
			return out;
		}
		throw std::logic_error("qboot_b200: decimal exponent search did not converge");
	}
}  // namespace qb::host
