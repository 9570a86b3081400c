// qboot_b200 façade -- the positive prefactor chi_j(x) of one polynomial inequality (call surface of
// include/qboot/scale_factor.hpp): degree bound, sample points x_k, sample scalings chi(x_k), bilinear bases q_m(x).
#ifndef QBOOT_B200_SCALE_FACTOR_HPP_
#define QBOOT_B200_SCALE_FACTOR_HPP_

#include <cstdint>

#include "qboot/algebra/matrix.hpp"
#include "qboot/algebra/polynomial.hpp"
#include "qboot/mp/real.hpp"

namespace qboot
{
	class ScaleFactor
	{
	public:
		ScaleFactor() = default;
		ScaleFactor(const ScaleFactor&) = default;
		ScaleFactor& operator=(const ScaleFactor&) = default;
		ScaleFactor(ScaleFactor&&) noexcept = default;
		ScaleFactor& operator=(ScaleFactor&&) noexcept = default;
		virtual ~ScaleFactor();
		[[nodiscard]] virtual uint32_t max_degree() const = 0;                               // D
		[[nodiscard]] virtual mp::real eval(const mp::real& v) const = 0;                    // chi(v), v >= 0
		[[nodiscard]] virtual algebra::Vector<mp::real> sample_scalings() const = 0;         // chi(x_0..x_D)
		[[nodiscard]] virtual mp::real sample_point(uint32_t k) const = 0;                   // x_k
		[[nodiscard]] virtual algebra::Vector<mp::real> sample_points() const = 0;           // x_0..x_D
		[[nodiscard]] virtual algebra::Vector<algebra::Polynomial> bilinear_bases() const = 0;  // q_0..q_{D/2}
	};
	// chi = 1, degree 0: constraints that do not depend on x (discrete sectors)
	class TrivialScale : public ScaleFactor
	{
	public:
		TrivialScale() = default;
		TrivialScale(const TrivialScale&) = delete;
		TrivialScale& operator=(const TrivialScale&) = delete;
		TrivialScale(TrivialScale&&) noexcept = default;
		TrivialScale& operator=(TrivialScale&&) noexcept = default;
		~TrivialScale() override;
		[[nodiscard]] uint32_t max_degree() const override { return 0; }
		[[nodiscard]] mp::real eval([[maybe_unused]] const mp::real& v) const override { return mp::real(1); }
		[[nodiscard]] algebra::Vector<mp::real> sample_scalings() const override { return algebra::Vector<mp::real>{mp::real(1)}; }
		[[nodiscard]] mp::real sample_point([[maybe_unused]] uint32_t k) const override { return mp::real(0); }
		[[nodiscard]] algebra::Vector<mp::real> sample_points() const override { return algebra::Vector<mp::real>{mp::real(1)}; }
		[[nodiscard]] algebra::Vector<algebra::Polynomial> bilinear_bases() const override
		{
			return algebra::Vector<algebra::Polynomial>{algebra::Polynomial(mp::real(1))};
		}
	};
}  // namespace qboot

#endif  // QBOOT_B200_SCALE_FACTOR_HPP_
