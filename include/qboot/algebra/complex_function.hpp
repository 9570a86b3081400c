// qboot_b200 façade -- block-derivative tables as the reference hands them out
// (include/qboot/algebra/complex_function.hpp:17-216).
//
// A ComplexFunction is the table of coefficients of (x - 1/2)^m y^n, m + 2n <= lambda, z = x + sqrt(y), packed
// dy-major / dx-minor -- exactly the record order the CUDA engine writes (qb_gblock_batch), so a table fetched from the
// GPU becomes a ComplexFunction without any shuffling.  The arithmetic members (mul, proj) exist for API parity and
// follow the reference's loop order; the production path multiplies tables on the device (qb_fblock_batch, qb_pmp_*).
#ifndef QBOOT_B200_ALGEBRA_COMPLEX_FUNCTION_HPP_
#define QBOOT_B200_ALGEBRA_COMPLEX_FUNCTION_HPP_

#include <cstdint>
#include <utility>

#include "qboot/algebra/matrix.hpp"
#include "qboot/mp/real.hpp"

namespace qboot::algebra
{
	// behaviour under z -> 1 - z
	enum class FunctionSymmetry : uint32_t
	{
		Even = 0,
		Odd = 1,
		Mixed
	};
	inline constexpr bool _matches(FunctionSymmetry s, uint32_t dx) noexcept
	{
		return s == FunctionSymmetry::Mixed || (dx % 2) == uint32_t(s);
	}
	inline constexpr FunctionSymmetry _times(FunctionSymmetry s, FunctionSymmetry t) noexcept
	{
		if (s == FunctionSymmetry::Mixed || t == FunctionSymmetry::Mixed) return FunctionSymmetry::Mixed;
		return s == t ? FunctionSymmetry::Even : FunctionSymmetry::Odd;
	}
	inline constexpr FunctionSymmetry _plus(FunctionSymmetry s, FunctionSymmetry t) noexcept
	{
		return s == t ? s : FunctionSymmetry::Mixed;
	}
	// number of (m, n) with m + 2n <= lambda and m of the given parity
	inline constexpr uint32_t function_dimension(uint32_t lambda, FunctionSymmetry sym = FunctionSymmetry::Mixed) noexcept
	{
		const uint32_t h = sym == FunctionSymmetry::Even ? (lambda + 2) / 2 : (lambda + 1) / 2;
		return sym == FunctionSymmetry::Mixed ? (lambda + 2) * (lambda + 2) / 4 : h * (h + 1) / 2;
	}

	template <class Ring>
	class ComplexFunction
	{
		FunctionSymmetry sym_ = FunctionSymmetry::Mixed;
		uint32_t lambda_ = 0;
		Vector<Ring> c_;
		[[nodiscard]] uint32_t index(uint32_t dx, uint32_t dy) const noexcept
		{
			if (sym_ == FunctionSymmetry::Even) return (lambda_ / 2 * 2 + 3 - dy) * dy / 2 + dx / 2;
			if (sym_ == FunctionSymmetry::Odd) return ((lambda_ - 1) / 2 * 2 + 3 - dy) * dy / 2 + (dx - 1) / 2;
			return (lambda_ + 2 - dy) * dy + dx;
		}

	public:
		void _reset() &&
		{
			sym_ = FunctionSymmetry::Mixed;
			lambda_ = 0;
			std::move(c_)._reset();
		}
		ComplexFunction() = default;
		explicit ComplexFunction(uint32_t lambda, FunctionSymmetry sym = FunctionSymmetry::Mixed)
		    : sym_(sym), lambda_(lambda), c_(function_dimension(lambda, sym))
		{
		}
		// adopt a packed coefficient vector (the engine's record order)
		ComplexFunction(Vector<Ring>&& packed, uint32_t lambda, FunctionSymmetry sym) : sym_(sym), lambda_(lambda), c_(std::move(packed)) {}
		[[nodiscard]] FunctionSymmetry symmetry() const { return sym_; }
		[[nodiscard]] uint32_t lambda() const { return lambda_; }
		[[nodiscard]] uint32_t size() const { return c_.size(); }
		void swap(ComplexFunction& o) &
		{
			std::swap(sym_, o.sym_);
			std::swap(lambda_, o.lambda_);
			c_.swap(o.c_);
		}
		void negate() & { c_.negate(); }
		[[nodiscard]] bool iszero() const { return c_.iszero(); }
		[[nodiscard]] ComplexFunction clone() const { return ComplexFunction(c_.clone(), lambda_, sym_); }
		[[nodiscard]] Ring& at(uint32_t dx, uint32_t dy) & { return c_[index(dx, dy)]; }
		[[nodiscard]] const Ring& at(uint32_t dx, uint32_t dy) const& { return c_[index(dx, dy)]; }
		[[nodiscard]] Vector<Ring> flatten() && { return std::move(c_); }
		[[nodiscard]] const Vector<Ring>& _packed() const { return c_; }
		[[nodiscard]] auto abs() const { return c_.abs(); }
		[[nodiscard]] auto norm() const { return c_.norm(); }
		ComplexFunction& operator+=(const ComplexFunction& o) &
		{
			c_ += o.c_;
			return *this;
		}
		ComplexFunction& operator-=(const ComplexFunction& o) &
		{
			c_ -= o.c_;
			return *this;
		}
		template <class R>
		ComplexFunction& operator*=(const R& r) &
		{
			c_ *= r;
			return *this;
		}
		template <class R>
		ComplexFunction& operator/=(const R& r) &
		{
			c_ /= r;
			return *this;
		}
		friend ComplexFunction operator+(const ComplexFunction& a, const ComplexFunction& b)
		{
			ComplexFunction r = a.clone();
			r += b;
			return r;
		}
		friend ComplexFunction operator-(const ComplexFunction& a, const ComplexFunction& b)
		{
			ComplexFunction r = a.clone();
			r -= b;
			return r;
		}
		ComplexFunction operator-() const&
		{
			ComplexFunction r = clone();
			r.negate();
			return r;
		}
		template <class R>
		friend ComplexFunction mul_scalar(const R& r, const ComplexFunction& f)
		{
			return ComplexFunction(mul_scalar(r, f.c_), f.lambda_, f.sym_);
		}
		template <class R>
		friend ComplexFunction mul_scalar(const R& r, ComplexFunction&& f)
		{
			return mul_scalar(r, static_cast<const ComplexFunction&>(f));
		}
		// truncated product: (f g)(M, N) = sum f(m1, n1) g(M - m1, N - n1), m1 outer / n1 inner per output as the reference
		// iterates (include/qboot/algebra/complex_function.hpp:190-206)
		friend ComplexFunction mul(const ComplexFunction& f, const ComplexFunction& g)
		{
			ComplexFunction z(f.lambda_, _times(f.sym_, g.sym_));
			for (uint32_t dy1 = 0; dy1 <= f.lambda_ / 2; ++dy1)
				for (uint32_t dx1 = 0; dx1 + 2 * dy1 <= f.lambda_; ++dx1)
				{
					if (!_matches(f.sym_, dx1)) continue;
					for (uint32_t dy2 = 0; dy1 + dy2 <= f.lambda_ / 2; ++dy2)
						for (uint32_t dx2 = 0; dx1 + dx2 + 2 * (dy1 + dy2) <= f.lambda_; ++dx2)
						{
							if (!_matches(g.sym_, dx2)) continue;
							z.at(dx1 + dx2, dy1 + dy2) += mul(f.at(dx1, dy1), g.at(dx2, dy2));
						}
				}
			return z;
		}
		// keep the coefficients whose dx has parity sym
		[[nodiscard]] ComplexFunction proj(FunctionSymmetry sym) &&
		{
			if (sym_ == sym) return std::move(*this);
			ComplexFunction z(lambda_, sym);
			if (sym_ == FunctionSymmetry::Mixed)
				for (uint32_t dy = 0; dy <= lambda_ / 2; ++dy)
					for (uint32_t dx = 0; dx + 2 * dy <= lambda_; ++dx)
						if (_matches(sym, dx)) z.at(dx, dy).swap(at(dx, dy));
			return z;
		}
	};
}  // namespace qboot::algebra

#endif  // QBOOT_B200_ALGEBRA_COMPLEX_FUNCTION_HPP_
