// qboot_b200 façade -- dense real polynomials (call surface of include/qboot/algebra/polynomial.hpp).
//
// Bilinear bases of a ConformalScale are handed around as Polynomials and evaluated by Horner's rule with fma
// (reference: polynomial.hpp:66-83), which is the order kept here; interpolation (inverse Vandermonde,
// src/algebra/polynomial.cpp:182-193) serves the XML writer.
#ifndef QBOOT_B200_ALGEBRA_POLYNOMIAL_HPP_
#define QBOOT_B200_ALGEBRA_POLYNOMIAL_HPP_

#include <cassert>
#include <cstdint>
#include <initializer_list>
#include <ostream>
#include <string>
#include <utility>

#include "qboot/algebra/matrix.hpp"
#include "qboot/mp/real.hpp"

namespace qboot::algebra
{
	// sum_i c[i] x^i; the zero polynomial has no coefficients, otherwise the leading one is nonzero
	class Polynomial
	{
		Vector<mp::real> c_;
		void trim()
		{
			uint32_t n = c_.size();
			while (n > 0 && c_[n - 1].iszero()) --n;
			if (n == c_.size()) return;
			Vector<mp::real> t(n);
			for (uint32_t i = 0; i < n; ++i) t[i] = c_[i];
			c_ = std::move(t);
		}

	public:
		void _reset() && { std::move(c_)._reset(); }
		Polynomial() = default;
		Polynomial(Polynomial&&) noexcept = default;
		Polynomial& operator=(Polynomial&&) & noexcept = default;
		Polynomial(const Polynomial&) = delete;
		Polynomial& operator=(const Polynomial&) = delete;
		~Polynomial() = default;
		// x^degree
		explicit Polynomial(uint32_t degree) : c_(degree + 1) { c_[degree] = mp::real(1); }
		explicit Polynomial(const mp::real& constant)
		{
			if (!constant.iszero()) c_ = Vector<mp::real>{constant};
		}
		Polynomial(const mp::real& lead, uint32_t degree) : c_(degree + 1) { c_[degree] = lead; }
		explicit Polynomial(Vector<mp::real>&& coeffs) : c_(std::move(coeffs)) { trim(); }
		Polynomial(std::initializer_list<mp::real> coeffs) : c_(coeffs) { trim(); }
		[[nodiscard]] const mp::real& at(uint32_t p) const { return c_.at(p); }
		const mp::real& operator[](uint32_t p) const { return c_[p]; }
		[[nodiscard]] const mp::real* begin() const& noexcept { return c_.begin(); }
		[[nodiscard]] const mp::real* end() const& noexcept { return c_.end(); }
		[[nodiscard]] bool iszero() const noexcept { return c_.size() == 0; }
		[[nodiscard]] int32_t degree() const noexcept { return int32_t(c_.size()) - 1; }
		[[nodiscard]] mp::real norm() const { return c_.norm(); }
		[[nodiscard]] mp::real abs() const { return mp::sqrt(norm()); }
		// Horner: real argument -> fma(s, x, c[i]); integer argument -> s *= x; s += c[i]   (two roundings, as mpfr_mul_ui + mpfr_add)
		template <class R>
		[[nodiscard]] mp::real eval(const R& x) const
		{
			if (iszero()) return mp::real{};
			const uint32_t d = uint32_t(degree());
			mp::real s = c_[d];
			for (uint32_t i = d; i-- > 0;)
			{
				if constexpr (std::is_same_v<R, mp::real>)
					mp::fma(s, s, x, c_[i]);
				else
				{
					s *= x;
					s += c_[i];
				}
			}
			return s;
		}
		[[nodiscard]] Polynomial clone() const
		{
			Polynomial p;
			p.c_ = c_.clone();
			return p;
		}
		void swap(Polynomial& o) & { c_.swap(o.c_); }
		void negate() & { c_.negate(); }
		void derivate() &
		{
			if (c_.size() <= 1)
			{
				c_ = {};
				return;
			}
			Vector<mp::real> t(c_.size() - 1);
			for (uint32_t i = 1; i < c_.size(); ++i) t[i - 1] = c_[i] * i;
			c_ = std::move(t);
		}
		Polynomial derivative() const
		{
			Polynomial t = clone();
			t.derivate();
			return t;
		}
		// this *= (a + x)
		void _mul_linear(const mp::real& a) &
		{
			if (iszero()) return;
			Vector<mp::real> t(c_.size() + 1);
			t[c_.size()] = c_[c_.size() - 1];
			for (uint32_t i = c_.size() - 1; i > 0; --i) mp::fma(t[i], c_[i], a, c_[i - 1]);
			t[0] = c_[0] * a;
			c_ = std::move(t);
		}
		Polynomial& operator+=(const Polynomial& p) &
		{
			if (p.c_.size() > c_.size())
			{
				Vector<mp::real> t(p.c_.size());
				for (uint32_t i = 0; i < c_.size(); ++i) t[i] = c_[i];
				c_ = std::move(t);
			}
			for (uint32_t i = 0; i < p.c_.size(); ++i) c_[i] += p.c_[i];
			trim();
			return *this;
		}
		Polynomial& operator-=(const Polynomial& p) &
		{
			if (p.c_.size() > c_.size())
			{
				Vector<mp::real> t(p.c_.size());
				for (uint32_t i = 0; i < c_.size(); ++i) t[i] = c_[i];
				c_ = std::move(t);
			}
			for (uint32_t i = 0; i < p.c_.size(); ++i) c_[i] -= p.c_[i];
			trim();
			return *this;
		}
		template <class R>
		Polynomial& operator*=(const R& c) &
		{
			c_ *= c;
			trim();
			return *this;
		}
		template <class R>
		Polynomial& operator/=(const R& c) &
		{
			c_ /= c;
			return *this;
		}
		Polynomial operator+() const& { return clone(); }
		Polynomial operator-() const&
		{
			Polynomial q = clone();
			q.negate();
			return q;
		}
		friend Polynomial mul(const Polynomial& p, const Polynomial& q)
		{
			if (p.iszero() || q.iszero()) return Polynomial{};
			Vector<mp::real> t(p.c_.size() + q.c_.size() - 1);
			for (uint32_t i = 0; i < p.c_.size(); ++i)
				for (uint32_t j = 0; j < q.c_.size(); ++j) t[i + j] += mp::mul(p.c_[i], q.c_[j]);
			return Polynomial(std::move(t));
		}
		friend Polynomial operator*(const Polynomial& p, const Polynomial& q) { return mul(p, q); }
		template <class R>
		friend Polynomial mul_scalar(const R& c, const Polynomial& p)
		{
			Polynomial r = p.clone();
			r *= c;
			return r;
		}
		friend Polynomial operator+(const Polynomial& p, const Polynomial& q)
		{
			Polynomial r = p.clone();
			r += q;
			return r;
		}
		friend Polynomial operator-(const Polynomial& p, const Polynomial& q)
		{
			Polynomial r = p.clone();
			r -= q;
			return r;
		}
		template <class R>
		friend Polynomial operator/(const Polynomial& p, const R& c)
		{
			Polynomial r = p.clone();
			r /= c;
			return r;
		}
		friend bool operator==(const Polynomial& p, const Polynomial& q) { return p.c_ == q.c_; }
		friend bool operator!=(const Polynomial& p, const Polynomial& q) { return !(p == q); }
		[[nodiscard]] std::string str(char var = 'x') const;
		friend std::ostream& operator<<(std::ostream& out, const Polynomial& p) { return out << p.str(); }
	};

	inline Polynomial to_pol(Vector<mp::real>* coeffs) { return Polynomial(coeffs->clone()); }
	// (deg + 1) x (deg + 1) matrix A with coefficients = A . values: the inverse Vandermonde of the points
	Matrix<mp::real> interpolation_matrix(const Vector<mp::real>& points);
	// coefficients of the polynomial through (points[i], vals[i])
	Polynomial polynomial_interpolate(const Vector<mp::real>& vals, const Matrix<mp::real>& interpolation_matrix);
	inline Polynomial polynomial_interpolate(const Vector<mp::real>& vals, const Vector<mp::real>& points)
	{
		return polynomial_interpolate(vals, interpolation_matrix(points));
	}
	template <class P>
	Vector<mp::real> evals(const P& v, const Vector<mp::real>& xs)
	{
		Vector<mp::real> r(xs.size());
		for (uint32_t i = 0; i < xs.size(); ++i) r[i] = v.eval(xs[i]);
		return r;
	}
}  // namespace qboot::algebra

#endif  // QBOOT_B200_ALGEBRA_POLYNOMIAL_HPP_
