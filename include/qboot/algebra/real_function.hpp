// qboot_b200 façade -- truncated Taylor series in one variable (call surface of include/qboot/algebra/real_function.hpp).
// The rho-series of a block lives on the device; this container only carries diagonal derivative vectors across the API
// (Context::expand_off_diagonal, gBlock_real) and the rho -> z conversion matrix read back from the engine.
#ifndef QBOOT_B200_ALGEBRA_REAL_FUNCTION_HPP_
#define QBOOT_B200_ALGEBRA_REAL_FUNCTION_HPP_

#include <cstdint>
#include <utility>

#include "qboot/algebra/matrix.hpp"
#include "qboot/mp/real.hpp"

namespace qboot::algebra
{
	// sum_{k <= lambda} at(k) (x - x0)^k
	template <class Ring>
	class RealFunction
	{
		uint32_t lambda_ = 0;
		Vector<Ring> c_;

	public:
		void _reset() &&
		{
			lambda_ = 0;
			std::move(c_)._reset();
		}
		RealFunction() = default;
		explicit RealFunction(uint32_t lambda) : lambda_(lambda), c_(lambda + 1) {}
		RealFunction(Vector<Ring>&& coeffs, uint32_t lambda) : lambda_(lambda), c_(std::move(coeffs)) {}
		[[nodiscard]] uint32_t lambda() const { return lambda_; }
		[[nodiscard]] uint32_t size() const { return c_.size(); }
		[[nodiscard]] Ring& at(uint32_t k) & { return c_[k]; }
		[[nodiscard]] const Ring& at(uint32_t k) const& { return c_[k]; }
		[[nodiscard]] RealFunction clone() const { return RealFunction(c_.clone(), lambda_); }
		[[nodiscard]] Vector<Ring> flatten() && { return std::move(c_); }
		[[nodiscard]] const Vector<Ring>& _coeffs() const { return c_; }
		void swap(RealFunction& o) &
		{
			std::swap(lambda_, o.lambda_);
			c_.swap(o.c_);
		}
		void negate() & { c_.negate(); }
		[[nodiscard]] bool iszero() const { return c_.iszero(); }
		// truncated product, k1 outer / k2 inner (include/qboot/algebra/real_function.hpp:121-128)
		friend RealFunction mul(const RealFunction& x, const RealFunction& y)
		{
			RealFunction z(x.lambda_);
			for (uint32_t k1 = 0; k1 <= x.lambda_; ++k1)
				for (uint32_t k2 = 0; k1 + k2 <= x.lambda_; ++k2) z.at(k1 + k2) += mul(x.at(k1), y.at(k2));
			return z;
		}
	};
	// the fixed (lambda+1)^2 matrix taking rho-Taylor coefficients to z-Taylor coefficients (Context::rho_to_z()); the
	// engine builds it once per Context (qb_context_init) and this class holds the copy read back with qb_context_get
	class RealConverter
	{
		uint32_t lambda_ = 0;
		Matrix<mp::real> mat_;

	public:
		RealConverter() = default;
		RealConverter(Matrix<mp::real>&& mat, uint32_t lambda) : lambda_(lambda), mat_(std::move(mat)) {}
		[[nodiscard]] uint32_t lambda() const { return lambda_; }
		[[nodiscard]] const Matrix<mp::real>& matrix() const { return mat_; }
		[[nodiscard]] uint32_t _total_memory() const { return mat_.row() * mat_.column(); }
		// vector x matrix (include/qboot/algebra/real_function.hpp:208-213)
		[[nodiscard]] RealFunction<mp::real> convert(const RealFunction<mp::real>& f) const
		{
			return RealFunction<mp::real>(dot(f._coeffs(), mat_), lambda_);
		}
	};
}  // namespace qboot::algebra

#endif  // QBOOT_B200_ALGEBRA_REAL_FUNCTION_HPP_
