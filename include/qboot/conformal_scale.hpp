// qboot_b200 façade -- the scale factor of a continuous sector (call surface of include/qboot/conformal_scale.hpp).
//
//   chi(Delta) = (4 rho*)^Delta / prod_i (Delta - p_i)   [times (x + ub)^-D for a finite range [lb, ub)]
// with the poles p_i of the block in Delta, the sample points x_k = pi^2 (4k - 1)^2 / (-64 (D + 1) ln rho*) and, for
// half-infinite ranges, bilinear bases orthonormal for the weight chi on x >= 0 (moments through E_1, Cholesky of the
// Hankel matrix; src/conformal_scale.cpp:11-82,136-187).  Everything here runs on the host once per constraint with
// the reference's operations in the reference's order; the sample Deltas it produces are the keys of the block tables
// the GPU then evaluates.
#ifndef QBOOT_B200_CONFORMAL_SCALE_HPP_
#define QBOOT_B200_CONFORMAL_SCALE_HPP_

#include <cstdint>
#include <optional>
#include <utility>

#include "qboot/algebra/matrix.hpp"
#include "qboot/algebra/polynomial.hpp"
#include "qboot/mp/rational.hpp"
#include "qboot/mp/real.hpp"
#include "qboot/primary_op.hpp"
#include "qboot/scale_factor.hpp"

namespace qboot
{
	class Context;
	inline bool include_odd(const mp::real& d1, const mp::real& d2, const mp::real& d3, const mp::real& d4) { return d1 != d2 && d3 != d4; }
	inline bool include_odd(const mp::real& d12, const mp::real& d34) { return d12 != 0 && d34 != 0; }
	class ConformalScale : public ScaleFactor
	{
		bool odd_included_ = false;
		uint32_t spin_ = 0, lambda_ = 0;
		mp::rational epsilon_;
		mp::real rho_, gap_;
		std::optional<mp::real> end_;
		algebra::Vector<mp::real> poles_;  // 1 / (Delta - poles_[i]), largest first
		algebra::Vector<algebra::Polynomial> bilinear_bases_{};
		bool bases_ready_ = true;
		void _set_bilinear_bases() &;

	public:
		// The planner of BootstrapEquation::convert only needs the poles and the sample points before the GPU can start; the
		// bilinear bases (incomplete gamma functions, a Hankel-Cholesky factorisation: most of the constructor's time) are
		// then built on host threads underneath the GPU phase.  _finish_bases() must run before bilinear_bases() is used.
		struct _DeferBases
		{
		};
		ConformalScale(_DeferBases, const GeneralPrimaryOperator& op, const Context& context, bool include_odd);
		void _finish_bases() &
		{
			if (!bases_ready_) _set_bilinear_bases();
			bases_ready_ = true;
		}
		void _reset() &&
		{
			odd_included_ = false;
			spin_ = lambda_ = 0;
			std::move(epsilon_)._reset();
			std::move(rho_)._reset();
			std::move(gap_)._reset();
			end_.reset();
			std::move(poles_)._reset();
			std::move(bilinear_bases_)._reset();
		}
		ConformalScale(const GeneralPrimaryOperator& op, const Context& context, bool include_odd);
		// gap <= Delta < end (end empty: no upper limit); cutoff = number of poles kept
		ConformalScale(uint32_t cutoff, uint32_t spin, const Context& context, bool include_odd, const mp::real& gap,
		               const std::optional<mp::real>& end);
		// the same without a Context (lambda, epsilon and rho* = 3 - sqrt(8) given explicitly)
		ConformalScale(uint32_t cutoff, uint32_t spin, uint32_t lambda, const mp::rational& epsilon, const mp::real& rho,
		               bool include_odd, const mp::real& gap, const std::optional<mp::real>& end);
		ConformalScale(const ConformalScale&) = delete;
		ConformalScale& operator=(const ConformalScale&) = delete;
		ConformalScale(ConformalScale&&) noexcept = default;
		ConformalScale& operator=(ConformalScale&&) noexcept = default;
		~ConformalScale() override;
		[[nodiscard]] bool odd_included() const { return odd_included_; }
		[[nodiscard]] uint32_t max_degree() const override { return poles_.size() + lambda_; }
		[[nodiscard]] const algebra::Vector<mp::real>& get_poles() const& { return poles_; }
		[[nodiscard]] algebra::Vector<mp::real> get_poles() && { return std::move(poles_); }
		// chi as a function of Delta
		[[nodiscard]] mp::real eval_d(const mp::real& delta) const;
		// x in [0, inf) -> Delta (a Moebius map for a finite range) and back
		[[nodiscard]] mp::real get_delta(const mp::real& x) const;
		[[nodiscard]] mp::real get_x(const mp::real& delta) const;
		[[nodiscard]] mp::real eval(const mp::real& x) const override { return eval_d(get_delta(x)); }
		[[nodiscard]] algebra::Vector<mp::real> sample_scalings() const override;
		[[nodiscard]] algebra::Vector<algebra::Polynomial> bilinear_bases() const override { return bilinear_bases_.clone(); }
		[[nodiscard]] mp::real sample_point(uint32_t k) const override { return sample_point(max_degree(), k); }
		[[nodiscard]] algebra::Vector<mp::real> sample_points() const override { return sample_points(max_degree()); }
		// x_k = pi^2 (4k - 1)^2 / (-64 (degree + 1) ln(3 - sqrt 8))
		static mp::real sample_point(uint32_t degree, uint32_t k);
		static algebra::Vector<mp::real> sample_points(uint32_t degree);
	};
}  // namespace qboot

#endif  // QBOOT_B200_CONFORMAL_SCALE_HPP_
