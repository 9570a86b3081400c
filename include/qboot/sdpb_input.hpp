// qboot_b200 façade -- the SDPB input directory (call surface of include/qboot/sdpb_input.hpp).
//
// Files and their layout are the reference's (src/sdpb_input.cpp:67-137): blocks.0, bilinear_bases.0, objectives,
// free_var_matrix.j, primal_objective_c.j; one number per line at 3 + int(0.302 prec) significant digits in the
// reference's default-float notation.  The two large files of every constraint are printed ON THE DEVICE from the packed
// d_B / d_c arrays that never left HBM (qb_pmp_text): only the bytes of the files cross PCIe.  Constraints registered
// from host data (DualConstraint built by user code) are printed by the host formatter, same digits.
#ifndef QBOOT_B200_SDPB_INPUT_HPP_
#define QBOOT_B200_SDPB_INPUT_HPP_

#include <array>
#include <cstdint>
#include <memory>
#include <optional>
#include <ostream>
#include <utility>
#include <vector>

#include "qboot/algebra/matrix.hpp"
#include "qboot/mp/real.hpp"
#include "qboot/my_filesystem.hpp"
#include "qboot/task_queue.hpp"

namespace qboot
{
	namespace b200
	{
		class Engine;
		struct PackedProgram;  // the packed arrays of one create_input on the device
	}
	// Tr(A_p Y) + (B y)_p = c_p for p <-> (r, s <= r, k): one positivity constraint of the dual program
	class DualConstraint
	{
		uint32_t dim_{}, deg_{};
		mutable algebra::Matrix<mp::real> constraint_B_{};
		mutable algebra::Vector<mp::real> constraint_c_{};
		std::array<algebra::Matrix<mp::real>, 2> bilinear_{};
		std::shared_ptr<b200::PackedProgram> dev_{};  // set when B and c live on the device
		int32_t slot_ = -1, cols_ = 0;
		void fetch() const;
		friend class SDPBInput;

	public:
		void _reset() &&
		{
			dim_ = deg_ = 0;
			std::move(constraint_B_)._reset();
			std::move(constraint_c_)._reset();
			for (auto&& x : bilinear_) std::move(x)._reset();
			dev_.reset();
			slot_ = -1;
		}
		DualConstraint() = default;
		DualConstraint(uint32_t dim, uint32_t deg, algebra::Matrix<mp::real>&& B, algebra::Vector<mp::real>&& c,
		               std::array<algebra::Matrix<mp::real>, 2>&& bilinear);
		// B and c stay on the device as entry `slot` of a packed program (cols = number of free variables)
		DualConstraint(uint32_t dim, uint32_t deg, std::shared_ptr<b200::PackedProgram> dev, int32_t slot, int32_t cols,
		               std::array<algebra::Matrix<mp::real>, 2>&& bilinear);
		[[nodiscard]] uint32_t dim() const { return dim_; }
		[[nodiscard]] uint32_t degree() const { return deg_; }
		[[nodiscard]] uint32_t schur_size() const { return (deg_ + 1) * dim_ * (dim_ + 1) / 2; }
		[[nodiscard]] const std::array<algebra::Matrix<mp::real>, 2>& bilinear() const { return bilinear_; }
		[[nodiscard]] uint32_t _total_memory() const
		{
			uint32_t sum = schur_size() * (uint32_t(dev_ ? cols_ : int32_t(constraint_B_.column())) + 1);
			for (const auto& q : bilinear_) sum += q.row() * q.column();
			return sum;
		}
		[[nodiscard]] uint32_t num_free() const { return dev_ ? uint32_t(cols_) : constraint_B_.column(); }
		// host copies (read back from the device on first use)
		[[nodiscard]] const algebra::Matrix<mp::real>& obj_B() const
		{
			fetch();
			return constraint_B_;
		}
		[[nodiscard]] const algebra::Vector<mp::real>& obj_c() const
		{
			fetch();
			return constraint_c_;
		}
	};

	// maximize b_0 + sum_n b_n y_n subject to the constraints
	class SDPBInput
	{
		mp::real constant_term_;
		algebra::Vector<mp::real> objectives_;
		std::unique_ptr<std::optional<DualConstraint>[]> constraints_;
		uint32_t num_constraints_;
		void write_all(const fs::path& root, uint32_t parallel, const std::unique_ptr<_event_base>& event) const;

	public:
		void _reset() &&
		{
			std::move(constant_term_)._reset();
			std::move(objectives_)._reset();
			constraints_.reset();
			num_constraints_ = 0;
		}
		SDPBInput(mp::real&& constant, algebra::Vector<mp::real>&& obj, uint32_t num_constraints);
		[[nodiscard]] uint32_t num_constraints() const { return num_constraints_; }
		[[nodiscard]] uint32_t _total_memory() const
		{
			uint32_t sum = 1 + objectives_.size();
			for (uint32_t i = 0; i < num_constraints_; ++i)
				if (constraints_[i].has_value()) sum += constraints_[i].value()._total_memory();
			return sum;
		}
		void register_constraint(uint32_t index, DualConstraint&& c) &;
		[[nodiscard]] const DualConstraint& constraint(uint32_t index) const { return constraints_[index].value(); }
		void write(const fs::path& root, uint32_t parallel = 1, const std::unique_ptr<_event_base>& event = {}) const&;
		void write(const fs::path& root, uint32_t parallel = 1, const std::unique_ptr<_event_base>& event = {}) &&;
	};
}  // namespace qboot

#endif  // QBOOT_B200_SDPB_INPUT_HPP_
