// qboot_b200 façade -- SDPB-1 style XML input (call surface of include/qboot/xml_input.hpp, src/xml_input.cpp:102-138).
// Legacy output: explicit polynomial coefficients instead of sample values.  Host side only; the polynomials come from
// interpolating the sample values of a program (PolynomialProgram::create_xml).
#ifndef QBOOT_B200_XML_INPUT_HPP_
#define QBOOT_B200_XML_INPUT_HPP_

#include <cstdint>
#include <memory>
#include <optional>
#include <utility>

#include "qboot/algebra/matrix.hpp"
#include "qboot/algebra/polynomial.hpp"
#include "qboot/mp/real.hpp"
#include "qboot/my_filesystem.hpp"

namespace qboot
{
	// one polynomial vector matrix: M(r, c)[0] is the target, [1 + m] multiplies the m-th free variable
	class PVM
	{
		uint32_t dim_{}, deg_{}, num_of_vars_{};
		algebra::Matrix<algebra::Vector<algebra::Polynomial>> M_{};
		algebra::Vector<mp::real> sample_points_{}, sample_scalings_{};
		algebra::Vector<algebra::Polynomial> bilinear_{};

	public:
		void _reset() &&
		{
			dim_ = deg_ = num_of_vars_ = 0;
			std::move(M_)._reset();
			std::move(sample_points_)._reset();
			std::move(sample_scalings_)._reset();
			std::move(bilinear_)._reset();
		}
		PVM() = default;
		PVM(algebra::Matrix<algebra::Vector<algebra::Polynomial>>&& M, algebra::Vector<mp::real>&& x, algebra::Vector<mp::real>&& scale,
		    algebra::Vector<algebra::Polynomial>&& bilinear)
		    : dim_(M.row()), deg_(x.size() - 1), num_of_vars_(M.at(0, 0).size() - 1), M_(std::move(M)), sample_points_(std::move(x)),
		      sample_scalings_(std::move(scale)), bilinear_(std::move(bilinear))
		{
		}
		[[nodiscard]] uint32_t dim() const { return dim_; }
		[[nodiscard]] uint32_t deg() const { return deg_; }
		[[nodiscard]] uint32_t num_of_vars() const { return num_of_vars_; }
		[[nodiscard]] const algebra::Vector<mp::real>& sample_points() const { return sample_points_; }
		[[nodiscard]] const algebra::Vector<mp::real>& sample_scalings() const { return sample_scalings_; }
		[[nodiscard]] const algebra::Matrix<algebra::Vector<algebra::Polynomial>>& matrices() const { return M_; }
		[[nodiscard]] const algebra::Vector<algebra::Polynomial>& bilinear() const { return bilinear_; }
	};
	class XMLInput
	{
		algebra::Vector<mp::real> objectives_;  // [0] the constant, [1 + m] the free variables
		std::unique_ptr<std::optional<PVM>[]> constraints_;
		uint32_t num_constraints_;

	public:
		void _reset() &&
		{
			std::move(objectives_)._reset();
			constraints_.reset();
			num_constraints_ = 0;
		}
		XMLInput(mp::real&& constant, algebra::Vector<mp::real>&& obj, uint32_t num_constraints);
		[[nodiscard]] uint32_t num_constraints() const { return num_constraints_; }
		void register_constraint(uint32_t index, PVM&& c) &;
		void write(const fs::path& path) const;
	};
}  // namespace qboot

#endif  // QBOOT_B200_XML_INPUT_HPP_
