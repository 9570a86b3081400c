// qboot_b200 façade -- operators of a CFT as the model description names them (call surface of the reference's
// include/qboot/primary_op.hpp).  Pure value types on the host; the engine receives (delta record, spin) pairs and redoes
// the divergence test and the shift on its own records (qboot_b200/csrc/engine.cuh, gblock_plan).
#ifndef QBOOT_B200_PRIMARY_OP_HPP_
#define QBOOT_B200_PRIMARY_OP_HPP_

#include <cassert>
#include <cstdint>
#include <optional>
#include <string>
#include <tuple>
#include <utility>

#include "qboot/mp/rational.hpp"
#include "qboot/mp/real.hpp"

namespace qboot
{
	class Context;
	// lowest dimension unitarity allows: epsilon for a scalar, spin + 2 epsilon otherwise
	inline mp::rational unitarity_bound(const mp::rational& epsilon, uint32_t spin)
	{
		return spin ? mp::rational(int64_t(spin)) + 2 * epsilon : epsilon;
	}
	namespace detail
	{
		// 1, 2, 3, ...: the values at which a linear factor of the recursion's leading polynomial vanishes
		inline bool natural(const mp::real& x) { return x > 0 && mp::isinteger(x); }
	}  // namespace detail

	// An operator of fixed scaling dimension: (Delta, spin) in the theory with dimension 2 epsilon + 2.
	class PrimaryOperator
	{
		struct Quantum
		{
			mp::real dim{};
			uint32_t spin = 0;
		};
		Quantum q_{};
		mp::rational eps_;
		[[nodiscard]] auto _order() const { return std::tie(q_.spin, q_.dim); }  // ordered by spin, then by dimension

	public:
		// the four ways to name one: unit operator, unitarity bound of a spin, explicit dimension -- each with a Context or
		// with epsilon itself
		explicit PrimaryOperator(const Context& c);
		PrimaryOperator(uint32_t spin, const Context& c);
		PrimaryOperator(const mp::real& delta, uint32_t spin, const Context& c);
		explicit PrimaryOperator(const mp::rational& epsilon) : eps_(epsilon) {}
		PrimaryOperator(uint32_t spin, const mp::rational& epsilon) : q_{mp::real(unitarity_bound(epsilon, spin)), spin}, eps_(epsilon) {}
		PrimaryOperator(const mp::real& delta, uint32_t spin, const mp::rational& epsilon) : q_{delta, spin}, eps_(epsilon) {}
		void _reset() &&
		{
			std::move(q_.dim)._reset();
			std::move(eps_)._reset();
			q_.spin = 0;
		}
		[[nodiscard]] uint32_t spin() const { return q_.spin; }
		[[nodiscard]] const mp::real& delta() const { return q_.dim; }
		[[nodiscard]] const mp::rational& epsilon() const { return eps_; }
		// eigenvalues of the two Casimir operators
		[[nodiscard]] mp::real quadratic_casimir() const
		{
			const mp::rational orbital = (2 * eps_ + q_.spin) * q_.spin / 2;
			return (q_.dim / 2 - eps_ - 1) * q_.dim + orbital;
		}
		[[nodiscard]] mp::real quartic_casimir() const
		{
			const mp::real a = q_.dim - 1;
			return q_.spin * (q_.spin + 2 * eps_) * a * (a - 2 * eps_);
		}
		// Does the series recursion divide by zero at some order n > 0?  Three families of dimensions do; the table of such an
		// operator is the average of two tables at Delta (1 +- small) (src/hor_recursion.cpp:21-28 in the reference).
		[[nodiscard]] bool is_divergent_hor() const
		{
			const mp::real& d = q_.dim;
			const int32_t l = int32_t(q_.spin);
			return detail::natural((1 - l) - d) || detail::natural(2 + 2 * eps_ - 2 * d) || (l > 0 && detail::natural((1 + l) + 2 * eps_ - d));
		}
		// the neighbour used for that average: Delta (1 + small); for the unit operator `small` itself
		[[nodiscard]] PrimaryOperator get_shifted(const mp::real& small) const
		{
			if (q_.dim == 0) return PrimaryOperator(small, q_.spin, eps_);
			return PrimaryOperator(q_.dim * (1 + small), q_.spin, eps_);
		}
		[[nodiscard]] std::string str() const;
		friend bool operator==(const PrimaryOperator& a, const PrimaryOperator& b) { return a._order() == b._order(); }
		friend bool operator!=(const PrimaryOperator& a, const PrimaryOperator& b) { return !(a == b); }
		friend bool operator<(const PrimaryOperator& a, const PrimaryOperator& b) { return a._order() < b._order(); }
		friend bool operator>(const PrimaryOperator& a, const PrimaryOperator& b) { return b < a; }
		friend bool operator<=(const PrimaryOperator& a, const PrimaryOperator& b) { return !(b < a); }
		friend bool operator>=(const PrimaryOperator& a, const PrimaryOperator& b) { return !(a < b); }
	};

	// An operator of given spin whose dimension ranges over [lower, upper) -- upper absent = unbounded -- together with the
	// number of poles its scale factor keeps (a continuous sector's entry).
	class GeneralPrimaryOperator
	{
		struct Range
		{
			mp::real lower;
			std::optional<mp::real> upper{};
		};
		Range range_;
		mp::rational eps_;
		uint32_t spin_ = 0, poles_kept_ = 0;

	public:
		// from the unitarity bound upwards
		GeneralPrimaryOperator(uint32_t spin, uint32_t num_poles, const mp::rational& epsilon)
		    : range_{mp::real(unitarity_bound(epsilon, spin)), {}}, eps_(epsilon), spin_(spin), poles_kept_(num_poles)
		{
		}
		GeneralPrimaryOperator(uint32_t spin, uint32_t num_poles, const mp::rational& epsilon, const mp::real& lb,
		                       const std::optional<mp::real>& ub = {})
		    : range_{lb, ub}, eps_(epsilon), spin_(spin), poles_kept_(num_poles)
		{
		}
		void _reset() &&
		{
			std::move(range_.lower)._reset();
			range_.upper.reset();
			std::move(eps_)._reset();
			spin_ = poles_kept_ = 0;
		}
		[[nodiscard]] uint32_t spin() const { return spin_; }
		[[nodiscard]] uint32_t num_poles() const { return poles_kept_; }
		[[nodiscard]] const mp::rational& epsilon() const { return eps_; }
		[[nodiscard]] bool is_finite() const noexcept { return range_.upper.has_value(); }
		[[nodiscard]] const mp::real& lower_bound() const { return range_.lower; }
		[[nodiscard]] const mp::real& upper_bound() const { return range_.upper.value(); }
		[[nodiscard]] const std::optional<mp::real>& upper_bound_safe() const { return range_.upper; }
		// the member of the family at one dimension inside the range
		[[nodiscard]] PrimaryOperator fix_delta(const mp::real& delta) const
		{
			assert(range_.lower <= delta && (!is_finite() || delta < upper_bound()));
			return PrimaryOperator(delta, spin_, eps_);
		}
		[[nodiscard]] std::string str() const;
	};
}  // namespace qboot

#endif  // QBOOT_B200_PRIMARY_OP_HPP_
