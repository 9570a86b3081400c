// qboot_b200 façade -- the positive prefactor chi_j(x) of one polynomial inequality (call surface of the reference's
// include/qboot/scale_factor.hpp).  User code may derive its own; the planner only needs the six queries below and calls
// them from host threads while the GPU works on the tables (qboot_b200/host/facade_program.cpp, Plan).
#ifndef QBOOT_B200_SCALE_FACTOR_HPP_
#define QBOOT_B200_SCALE_FACTOR_HPP_

#include <cstdint>

#include "qboot/algebra/matrix.hpp"
#include "qboot/algebra/polynomial.hpp"
#include "qboot/mp/real.hpp"

namespace qboot
{
	// Queries, with D = max_degree():
	//   eval(v)            chi(v) for v >= 0
	//   sample_point(k)    x_k,  k = 0..D          sample_points()    all of them
	//   sample_scalings()  chi(x_0) .. chi(x_D)
	//   bilinear_bases()   q_0 .. q_{D/2}, orthogonal under chi
	class ScaleFactor
	{
	protected:
		using Reals = algebra::Vector<mp::real>;
		using Polys = algebra::Vector<algebra::Polynomial>;

	public:
		ScaleFactor() = default;
		virtual ~ScaleFactor();
		ScaleFactor(const ScaleFactor&) = default;
		ScaleFactor(ScaleFactor&&) noexcept = default;
		ScaleFactor& operator=(const ScaleFactor&) = default;
		ScaleFactor& operator=(ScaleFactor&&) noexcept = default;

		[[nodiscard]] virtual uint32_t max_degree() const = 0;
		[[nodiscard]] virtual mp::real eval(const mp::real& v) const = 0;
		[[nodiscard]] virtual mp::real sample_point(uint32_t k) const = 0;
		[[nodiscard]] virtual Reals sample_points() const = 0;
		[[nodiscard]] virtual Reals sample_scalings() const = 0;
		[[nodiscard]] virtual Polys bilinear_bases() const = 0;
	};

	// chi = 1 with D = 0: the factor of a constraint that does not depend on x (discrete sectors).  Move-only, like every
	// scale factor the library hands out; bodies in qboot_b200/host/facade_math.cpp.
	class TrivialScale final : public ScaleFactor
	{
	public:
		TrivialScale() = default;
		~TrivialScale() override;
		TrivialScale(TrivialScale&&) noexcept = default;
		TrivialScale& operator=(TrivialScale&&) noexcept = default;
		TrivialScale(const TrivialScale&) = delete;
		TrivialScale& operator=(const TrivialScale&) = delete;

		[[nodiscard]] uint32_t max_degree() const override;
		[[nodiscard]] mp::real eval(const mp::real& v) const override;
		[[nodiscard]] mp::real sample_point(uint32_t k) const override;
		[[nodiscard]] Reals sample_points() const override;
		[[nodiscard]] Reals sample_scalings() const override;
		[[nodiscard]] Polys bilinear_bases() const override;
	};
}  // namespace qboot

#endif  // QBOOT_B200_SCALE_FACTOR_HPP_
