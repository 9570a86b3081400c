// qboot_b200 façade -- polynomial matrix programs (call surface of include/qboot/polynomial_program.hpp).
//
//   maximize  sum_n b[n] y[n] + b[N]   subject to   sum_n y[n] M_e[n] = M_e[N]   (equations, constants)
//                                                   sum_n y[n] M_j[n](x) >= M_j[N](x) for x >= 0   (inequalities)
// An inequality produced by BootstrapEquation::convert keeps its sample values ON THE DEVICE, as one constraint block
// of the engine's arena (qb_pmp_assemble); one built from host data is uploaded when the program is turned into SDPB
// input.  create_input then runs as a single pack kernel over all blocks (qb_pmp_pack): negation, re-indexing to
// (r, c <= r, k) rows and free-variable columns, and elimination of the pivoted variables with fma, in the reference's
// order (src/polynomial_program.cpp:194-221).  The normalisation equations themselves (add_equation, a handful of
// N-vectors) are handled on the host exactly as in src/polynomial_program.cpp:118-151.
#ifndef QBOOT_B200_POLYNOMIAL_PROGRAM_HPP_
#define QBOOT_B200_POLYNOMIAL_PROGRAM_HPP_

#include <cassert>
#include <cstdint>
#include <memory>
#include <array>
#include <optional>
#include <utility>
#include <vector>

#include "qboot/algebra/matrix.hpp"
#include "qboot/algebra/polynomial.hpp"
#include "qboot/mp/real.hpp"
#include "qboot/scale_factor.hpp"
#include "qboot/sdpb_input.hpp"
#include "qboot/task_queue.hpp"
#include "qboot/xml_input.hpp"

namespace qboot
{
	namespace b200
	{
		class Engine;
		struct DeviceBlock  // a constraint block resident in an engine's arena
		{
			std::shared_ptr<Engine> engine;
			int32_t block = -1;
			uint64_t epoch = 0;  // arena generation the block belongs to
		};
	}
	// One matrix inequality  sum_n y[n] M[n](x) >= M[N](x)  for x >= 0, held as sample values at x_0 .. x_D; M / chi is a
	// polynomial matrix of degree <= D.  After BootstrapEquation::convert the samples live in HBM (dev_) and the host copy
	// stays empty until somebody asks for it.
	class PolynomialInequality
	{
		using RealMatrix = algebra::Matrix<mp::real>;
		using Samples = algebra::Vector<RealMatrix>;  // one matrix per sample point
		uint32_t N_, sz_;
		std::unique_ptr<ScaleFactor> chi_;
		mutable algebra::Vector<Samples> mat_;  // mat_[n][k], sz x sz
		Samples target_{};
		b200::DeviceBlock dev_{};
		// bilinear bases at the sample points when convert has already evaluated them (underneath its GPU phase)
		std::optional<std::array<RealMatrix, 2>> bilinear_{};
		void fetch() const;  // device -> mat_
		[[nodiscard]] mp::real scale_at(uint32_t k) const { return chi_->eval(chi_->sample_point(k)); }
		[[nodiscard]] algebra::Matrix<algebra::Polynomial> as_polynomial(Samples&& vals);
		friend class BootstrapEquation;
		friend class PolynomialProgram;

	public:
		// scalar with a scale factor; scalar constant; matrix constant; matrix with a scale factor; device-resident samples
		// with zero target
		PolynomialInequality(uint32_t N, std::unique_ptr<ScaleFactor>&& scale, algebra::Vector<algebra::Vector<mp::real>>&& mat,
		                     algebra::Vector<mp::real>&& target);
		PolynomialInequality(uint32_t N, algebra::Vector<mp::real>&& mat, mp::real&& target);
		PolynomialInequality(uint32_t N, uint32_t sz, Samples&& mat, RealMatrix&& target);
		PolynomialInequality(uint32_t N, uint32_t sz, std::unique_ptr<ScaleFactor>&& scale, algebra::Vector<Samples>&& mat, Samples&& target);
		PolynomialInequality(uint32_t N, uint32_t sz, std::unique_ptr<ScaleFactor>&& scale, b200::DeviceBlock dev);
		~PolynomialInequality() = default;
		PolynomialInequality(PolynomialInequality&&) noexcept = default;
		PolynomialInequality& operator=(PolynomialInequality&&) noexcept = default;
		PolynomialInequality(const PolynomialInequality&) = delete;
		PolynomialInequality& operator=(const PolynomialInequality&) = delete;
		void _reset() &&;
		[[nodiscard]] uint32_t num_of_variables() const { return N_; }
		[[nodiscard]] uint32_t size() const { return sz_; }
		[[nodiscard]] uint32_t max_degree() const { return chi_->max_degree(); }
		[[nodiscard]] uint32_t _total_memory() const { return (N_ + 1) * (max_degree() + 1) * sz_ * sz_; }
		[[nodiscard]] bool _on_device() const { return dev_.block >= 0; }
		[[nodiscard]] algebra::Vector<mp::real> sample_points() const { return chi_->sample_points(); }
		[[nodiscard]] algebra::Vector<mp::real> sample_scalings() const { return chi_->sample_scalings(); }
		[[nodiscard]] algebra::Vector<algebra::Polynomial> bilinear_bases() const { return chi_->bilinear_bases(); }
		// Every accessor below CONSUMES what it returns, as in the reference.
		// M[n](x_k) and M[N](x_k), with the factor chi(x_k) ...
		[[nodiscard]] RealMatrix matrix_eval_with_scale(uint32_t n, uint32_t k) { return fetch(), std::move(mat_[n][k]); }
		[[nodiscard]] RealMatrix target_eval_with_scale(uint32_t k) { return std::move(target_[k]); }
		// ... and divided by it
		[[nodiscard]] RealMatrix matrix_eval_without_scale(uint32_t n, uint32_t k) { return matrix_eval_with_scale(n, k) / scale_at(k); }
		[[nodiscard]] RealMatrix target_eval_without_scale(uint32_t k) { return target_eval_with_scale(k) / scale_at(k); }
		// M[n] / chi and M[N] / chi as polynomial matrices (interpolation through the sample points)
		[[nodiscard]] algebra::Matrix<algebra::Polynomial> matrix_polynomial(uint32_t n) { return fetch(), as_polynomial(std::move(mat_[n])); }
		[[nodiscard]] algebra::Matrix<algebra::Polynomial> target_polynomial() { return as_polynomial(std::move(target_)); }
	};

	// maximize offset + objective . y  subject to equations (eliminated on the fly, one pivot each) and inequalities
	class PolynomialProgram
	{
		uint32_t N_;
		mp::real objective_offset_{};
		algebra::Vector<mp::real> objective_;
		std::vector<algebra::Vector<mp::real>> eq_rows_{};
		std::vector<mp::real> eq_rhs_{};
		std::vector<uint32_t> pivot_cols_{};  // y[pivot_cols_[e]] is fixed by equation e
		std::vector<uint32_t> free_cols_{};   // the others, ascending
		std::vector<std::optional<PolynomialInequality>> ineqs_{};
		std::shared_ptr<b200::Engine> engine_{};  // where create_input runs; taken from the first device-resident inequality
		// objective after elimination: (constant, coefficients of the free variables)
		std::pair<mp::real, algebra::Vector<mp::real>> eliminated_objective();

	public:
		explicit PolynomialProgram(uint32_t num_of_vars);
		void _reset() &&;
		[[nodiscard]] uint32_t num_of_variables() const { return N_; }
		[[nodiscard]] uint32_t num_of_inequalities() const { return uint32_t(ineqs_.size()); }
		[[nodiscard]] uint32_t _total_memory() const;
		[[nodiscard]] const mp::real& objective_constant() const { return objective_offset_; }
		[[nodiscard]] mp::real& objective_constant() { return objective_offset_; }
		[[nodiscard]] const algebra::Vector<mp::real>& objectives() const { return objective_; }
		void objectives(algebra::Vector<mp::real>&& obj) &;
		void _use_engine(std::shared_ptr<b200::Engine> e) & { engine_ = std::move(e); }
		// add sum_n y[n] vec[n] = target; equations must be independent; the order of calls fixes the output
		void add_equation(algebra::Vector<mp::real>&& vec, mp::real&& target) &;
		void add_inequality(std::optional<PolynomialInequality>&& ineq) &;
		// y (all N variables) from z (the free ones) through the equations
		algebra::Vector<mp::real> recover(const algebra::Vector<mp::real>& z);
		SDPBInput create_input(uint32_t parallel = 1, const std::unique_ptr<_event_base>& event = {}) &&;
		XMLInput create_xml(uint32_t parallel = 1, const std::unique_ptr<_event_base>& event = {}) &&;
	};
}  // namespace qboot

#endif  // QBOOT_B200_POLYNOMIAL_PROGRAM_HPP_
