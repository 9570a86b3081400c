// qboot_b200 façade -- SDPB-1 style XML input (call surface of the reference's include/qboot/xml_input.hpp; written by
// src/xml_input.cpp:102-138 there).  Legacy output: explicit polynomial coefficients instead of sample values.  Host side
// only; the polynomials come from interpolating the sample values of a program (PolynomialProgram::create_xml).
#ifndef QBOOT_B200_XML_INPUT_HPP_
#define QBOOT_B200_XML_INPUT_HPP_

#include <cstdint>
#include <optional>
#include <utility>
#include <vector>

#include "qboot/algebra/matrix.hpp"
#include "qboot/algebra/polynomial.hpp"
#include "qboot/mp/real.hpp"
#include "qboot/my_filesystem.hpp"

namespace qboot
{
	// One polynomial vector matrix.  Entry (r, c) is a vector of polynomials: [0] the target, [1 + m] the coefficient of the
	// m-th free variable.  Size, degree and variable count are read off the containers; nothing is cached.
	class PVM
	{
		using PolyVectors = algebra::Matrix<algebra::Vector<algebra::Polynomial>>;
		struct Payload
		{
			PolyVectors entries{};
			algebra::Vector<mp::real> points{}, scalings{};
			algebra::Vector<algebra::Polynomial> bases{};
		};
		Payload data_{};

	public:
		PVM() = default;
		PVM(PolyVectors&& M, algebra::Vector<mp::real>&& x, algebra::Vector<mp::real>&& scale, algebra::Vector<algebra::Polynomial>&& bilinear)
		{
			data_.entries = std::move(M);
			data_.points = std::move(x);
			data_.scalings = std::move(scale);
			data_.bases = std::move(bilinear);
		}
		// early release of the storage, the reference's idiom for consumed objects
		void _reset() && { data_ = Payload{}; }
		[[nodiscard]] uint32_t dim() const { return data_.entries.row(); }
		[[nodiscard]] uint32_t deg() const { return data_.points.size() == 0 ? 0 : data_.points.size() - 1; }
		[[nodiscard]] uint32_t num_of_vars() const { return dim() == 0 ? 0 : data_.entries.at(0, 0).size() - 1; }
		[[nodiscard]] const PolyVectors& matrices() const { return data_.entries; }
		[[nodiscard]] const algebra::Vector<mp::real>& sample_points() const { return data_.points; }
		[[nodiscard]] const algebra::Vector<mp::real>& sample_scalings() const { return data_.scalings; }
		[[nodiscard]] const algebra::Vector<algebra::Polynomial>& bilinear() const { return data_.bases; }
	};
	// The whole program: an objective vector ([0] the constant term, [1 + m] the free variables) and a slot per constraint,
	// filled in any order (PolynomialProgram::create_xml fills them from its host threads).
	class XMLInput
	{
		algebra::Vector<mp::real> objective_row_;
		std::vector<std::optional<PVM>> slots_;

	public:
		XMLInput(mp::real&& constant, algebra::Vector<mp::real>&& obj, uint32_t num_constraints);
		void _reset() &&
		{
			std::move(objective_row_)._reset();
			slots_.clear();
		}
		[[nodiscard]] uint32_t num_constraints() const { return uint32_t(slots_.size()); }
		void register_constraint(uint32_t index, PVM&& c) &;
		void write(const fs::path& path) const;
	};
}  // namespace qboot

#endif  // QBOOT_B200_XML_INPUT_HPP_
