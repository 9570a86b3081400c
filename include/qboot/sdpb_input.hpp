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
	// One positivity constraint of the dual program: Tr(A_p Y) + (B y)_p = c_p with p <-> (r, s <= r, k).  B and c live
	// either on the host (built by user code) or as entry `slot` of a packed program in HBM; the host copies appear on demand.
	class DualConstraint
	{
		using RealMatrix = algebra::Matrix<mp::real>;
		using Bases = std::array<RealMatrix, 2>;
		struct Shape
		{
			uint32_t dim = 0, deg = 0;
		};
		Shape shape_{};
		mutable RealMatrix host_B_{};
		mutable algebra::Vector<mp::real> host_c_{};
		Bases bilinear_{};
		std::shared_ptr<b200::PackedProgram> dev_{};
		int32_t slot_ = -1, cols_ = 0;
		void fetch() const;
		friend class SDPBInput;

	public:
		DualConstraint() = default;
		DualConstraint(uint32_t dim, uint32_t deg, RealMatrix&& B, algebra::Vector<mp::real>&& c, Bases&& bilinear);
		// device-resident B and c (cols = number of free variables)
		DualConstraint(uint32_t dim, uint32_t deg, std::shared_ptr<b200::PackedProgram> dev, int32_t slot, int32_t cols, Bases&& bilinear);
		void _reset() &&;
		[[nodiscard]] uint32_t dim() const { return shape_.dim; }
		[[nodiscard]] uint32_t degree() const { return shape_.deg; }
		// rows of B: one per (sample point, upper-triangle entry)
		[[nodiscard]] uint32_t schur_size() const { return (shape_.deg + 1) * shape_.dim * (shape_.dim + 1) / 2; }
		[[nodiscard]] uint32_t num_free() const { return dev_ ? uint32_t(cols_) : host_B_.column(); }
		[[nodiscard]] const Bases& bilinear() const { return bilinear_; }
		[[nodiscard]] uint32_t _total_memory() const;  // numbers held, the reference's progress measure
		[[nodiscard]] const RealMatrix& obj_B() const { return fetch(), host_B_; }
		[[nodiscard]] const algebra::Vector<mp::real>& obj_c() const { return fetch(), host_c_; }
	};

	// maximize offset + sum_n weights[n] y[n] subject to the constraints; slots are filled in any order
	class SDPBInput
	{
		mp::real offset_;
		algebra::Vector<mp::real> weights_;
		std::vector<std::optional<DualConstraint>> slots_;
		void write_all(const fs::path& root, uint32_t parallel, const std::unique_ptr<_event_base>& event) const;

	public:
		SDPBInput(mp::real&& constant, algebra::Vector<mp::real>&& obj, uint32_t num_constraints);
		void _reset() &&;
		[[nodiscard]] uint32_t num_constraints() const { return uint32_t(slots_.size()); }
		[[nodiscard]] uint32_t _total_memory() const;
		void register_constraint(uint32_t index, DualConstraint&& c) &;
		[[nodiscard]] const DualConstraint& constraint(uint32_t index) const { return slots_[index].value(); }
		// the directory SDPB reads; the rvalue form releases every constraint as soon as its files are written
		void write(const fs::path& root, uint32_t parallel = 1, const std::unique_ptr<_event_base>& event = {}) const&;
		void write(const fs::path& root, uint32_t parallel = 1, const std::unique_ptr<_event_base>& event = {}) &&;
	};
}  // namespace qboot

#endif  // QBOOT_B200_SDPB_INPUT_HPP_
