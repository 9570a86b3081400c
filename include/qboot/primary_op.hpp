// qboot_b200 façade -- operators of a CFT as the model description names them (call surface of
// include/qboot/primary_op.hpp).  Pure value types on the host; the engine receives (delta record, spin) pairs.
#ifndef QBOOT_B200_PRIMARY_OP_HPP_
#define QBOOT_B200_PRIMARY_OP_HPP_

#include <cassert>
#include <cstdint>
#include <optional>
#include <string>
#include <utility>

#include "qboot/mp/rational.hpp"
#include "qboot/mp/real.hpp"

namespace qboot
{
	class Context;
	inline bool _is_positive_integer(const mp::real& x) { return x > 0 && mp::isinteger(x); }
	// spin 0: epsilon; otherwise spin + 2 epsilon
	inline mp::rational unitarity_bound(const mp::rational& epsilon, uint32_t spin)
	{
		if (spin == 0) return epsilon;
		return mp::rational(int64_t(spin)) + 2 * epsilon;
	}
	// an operator of fixed scaling dimension
	class PrimaryOperator
	{
		mp::real delta_{};
		mp::rational epsilon_;
		uint32_t spin_ = 0;

	public:
		void _reset() &&
		{
			std::move(delta_)._reset();
			std::move(epsilon_)._reset();
			spin_ = 0;
		}
		explicit PrimaryOperator(const Context& c);                       // the unit operator
		explicit PrimaryOperator(const mp::rational& epsilon) : epsilon_(epsilon) {}
		PrimaryOperator(uint32_t spin, const Context& c);                 // on the unitarity bound
		PrimaryOperator(uint32_t spin, const mp::rational& epsilon) : delta_(unitarity_bound(epsilon, spin)), epsilon_(epsilon), spin_(spin) {}
		PrimaryOperator(const mp::real& delta, uint32_t spin, const Context& c);
		PrimaryOperator(const mp::real& delta, uint32_t spin, const mp::rational& epsilon) : delta_(delta), epsilon_(epsilon), spin_(spin) {}
		[[nodiscard]] const mp::real& delta() const { return delta_; }
		[[nodiscard]] uint32_t spin() const { return spin_; }
		[[nodiscard]] const mp::rational& epsilon() const { return epsilon_; }
		// delta (1 + small), or `small` itself for the unit operator (src/hor_recursion.cpp:21-28 evaluates divergent blocks there)
		[[nodiscard]] PrimaryOperator get_shifted(const mp::real& small) const
		{
			return PrimaryOperator(delta_ == 0 ? small : delta_ * (1 + small), spin_, epsilon_);
		}
		[[nodiscard]] mp::real quadratic_casimir() const
		{
			return (2 * epsilon_ + spin_) * spin_ / 2 + (delta_ / 2 - epsilon_ - 1) * delta_;
		}
		[[nodiscard]] mp::real quartic_casimir() const
		{
			return spin_ * (spin_ + 2 * epsilon_) * (delta_ - 1) * (delta_ - 1 - 2 * epsilon_);
		}
		// does the recursion for h divide by zero at some n > 0?  (one of the linear factors of p_0(n) vanishes)
		[[nodiscard]] bool is_divergent_hor() const
		{
			if (_is_positive_integer((1 - int32_t(spin_)) - delta_)) return true;
			if (_is_positive_integer(2 + 2 * epsilon_ - 2 * delta_)) return true;
			return spin_ > 0 && _is_positive_integer((1 + spin_) + 2 * epsilon_ - delta_);
		}
		[[nodiscard]] std::string str() const;
		friend bool operator==(const PrimaryOperator& p, const PrimaryOperator& q) { return p.spin_ == q.spin_ && p.delta_ == q.delta_; }
		friend bool operator!=(const PrimaryOperator& p, const PrimaryOperator& q) { return !(p == q); }
		friend bool operator<(const PrimaryOperator& p, const PrimaryOperator& q)
		{
			return p.spin_ != q.spin_ ? p.spin_ < q.spin_ : p.delta_ < q.delta_;
		}
		friend bool operator>(const PrimaryOperator& p, const PrimaryOperator& q) { return q < p; }
		friend bool operator>=(const PrimaryOperator& p, const PrimaryOperator& q) { return !(p < q); }
		friend bool operator<=(const PrimaryOperator& p, const PrimaryOperator& q) { return !(q < p); }
	};
	// an operator whose dimension ranges over [lb, ub) (ub empty = infinity), with the number of poles kept in its scale factor
	class GeneralPrimaryOperator
	{
		mp::rational epsilon_;
		mp::real from_;
		std::optional<mp::real> to_{};
		uint32_t spin_ = 0, num_poles_ = 0;

	public:
		void _reset() &&
		{
			std::move(epsilon_)._reset();
			std::move(from_)._reset();
			to_.reset();
			spin_ = num_poles_ = 0;
		}
		GeneralPrimaryOperator(uint32_t spin, uint32_t num_poles, const mp::rational& epsilon)
		    : epsilon_(epsilon), from_(unitarity_bound(epsilon, spin)), spin_(spin), num_poles_(num_poles)
		{
		}
		GeneralPrimaryOperator(uint32_t spin, uint32_t num_poles, const mp::rational& epsilon, const mp::real& lb,
		                       const std::optional<mp::real>& ub = {})
		    : epsilon_(epsilon), from_(lb), to_(ub), spin_(spin), num_poles_(num_poles)
		{
		}
		[[nodiscard]] uint32_t spin() const { return spin_; }
		[[nodiscard]] uint32_t num_poles() const { return num_poles_; }
		[[nodiscard]] const mp::rational& epsilon() const { return epsilon_; }
		[[nodiscard]] const mp::real& lower_bound() const { return from_; }
		[[nodiscard]] bool is_finite() const noexcept { return to_.has_value(); }
		[[nodiscard]] const mp::real& upper_bound() const { return to_.value(); }
		[[nodiscard]] const std::optional<mp::real>& upper_bound_safe() const { return to_; }
		[[nodiscard]] PrimaryOperator fix_delta(const mp::real& delta) const
		{
			assert(from_ <= delta && (!to_.has_value() || delta < to_.value()));
			return PrimaryOperator(delta, spin_, epsilon_);
		}
		[[nodiscard]] std::string str() const;
	};
}  // namespace qboot

#endif  // QBOOT_B200_PRIMARY_OP_HPP_
