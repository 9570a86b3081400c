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
	// sum_n y[n] M[n](x) >= M[N](x), sampled at x_0 .. x_D; M / chi is polynomial of degree <= D
	class PolynomialInequality
	{
		uint32_t N_, sz_;
		std::unique_ptr<ScaleFactor> chi_;
		// host copy of the sample values mat_[n][k] (sz x sz); empty while the values only exist on the device
		mutable algebra::Vector<algebra::Vector<algebra::Matrix<mp::real>>> mat_;
		algebra::Vector<algebra::Matrix<mp::real>> target_{};
		b200::DeviceBlock dev_{};
		// bilinear bases at the sample points when convert has already evaluated them (underneath its GPU phase)
		std::optional<std::array<algebra::Matrix<mp::real>, 2>> bilinear_{};
		friend class BootstrapEquation;
		void fetch() const;  // device -> mat_
		[[nodiscard]] algebra::Matrix<algebra::Polynomial> as_polynomial(algebra::Vector<algebra::Matrix<mp::real>>&& vals);
		friend class PolynomialProgram;

	public:
		void _reset() &&
		{
			N_ = sz_ = 0;
			chi_.reset();
			std::move(mat_)._reset();
			std::move(target_)._reset();
			dev_ = {};
		}
		PolynomialInequality(uint32_t N, std::unique_ptr<ScaleFactor>&& scale, algebra::Vector<algebra::Vector<mp::real>>&& mat,
		                     algebra::Vector<mp::real>&& target);
		PolynomialInequality(uint32_t N, algebra::Vector<mp::real>&& mat, mp::real&& target);
		PolynomialInequality(uint32_t N, uint32_t sz, algebra::Vector<algebra::Matrix<mp::real>>&& mat, algebra::Matrix<mp::real>&& target);
		PolynomialInequality(uint32_t N, uint32_t sz, std::unique_ptr<ScaleFactor>&& scale,
		                     algebra::Vector<algebra::Vector<algebra::Matrix<mp::real>>>&& mat,
		                     algebra::Vector<algebra::Matrix<mp::real>>&& target);
		// sample values resident on the device (zero target)
		PolynomialInequality(uint32_t N, uint32_t sz, std::unique_ptr<ScaleFactor>&& scale, b200::DeviceBlock dev);
		PolynomialInequality(const PolynomialInequality&) = delete;
		PolynomialInequality& operator=(const PolynomialInequality&) = delete;
		PolynomialInequality(PolynomialInequality&&) noexcept = default;
		PolynomialInequality& operator=(PolynomialInequality&&) noexcept = default;
		~PolynomialInequality() = default;
		[[nodiscard]] uint32_t num_of_variables() const { return N_; }
		[[nodiscard]] uint32_t _total_memory() const { return (N_ + 1) * (max_degree() + 1) * sz_ * sz_; }
		[[nodiscard]] uint32_t size() const { return sz_; }
		[[nodiscard]] uint32_t max_degree() const { return chi_->max_degree(); }
		[[nodiscard]] bool _on_device() const { return dev_.block >= 0; }
		[[nodiscard]] algebra::Vector<algebra::Polynomial> bilinear_bases() const { return chi_->bilinear_bases(); }
		[[nodiscard]] algebra::Vector<mp::real> sample_points() const { return chi_->sample_points(); }
		[[nodiscard]] algebra::Vector<mp::real> sample_scalings() const { return chi_->sample_scalings(); }
		// M[n] / chi and M[N] / chi as polynomial matrices (interpolation through the sample points)
		[[nodiscard]] algebra::Matrix<algebra::Polynomial> matrix_polynomial(uint32_t n)
		{
			fetch();
			return as_polynomial(std::move(mat_[n]));
		}
		[[nodiscard]] algebra::Matrix<algebra::Polynomial> target_polynomial() { return as_polynomial(std::move(target_)); }
		// M[n](x_k), M[N](x_k): consumed by the call, as in the reference
		[[nodiscard]] algebra::Matrix<mp::real> matrix_eval_with_scale(uint32_t n, uint32_t k)
		{
			fetch();
			return std::move(mat_[n][k]);
		}
		[[nodiscard]] algebra::Matrix<mp::real> target_eval_with_scale(uint32_t k) { return std::move(target_[k]); }
		[[nodiscard]] algebra::Matrix<mp::real> matrix_eval_without_scale(uint32_t n, uint32_t k)
		{
			fetch();
			return std::move(mat_[n][k]) / chi_->eval(chi_->sample_point(k));
		}
		[[nodiscard]] algebra::Matrix<mp::real> target_eval_without_scale(uint32_t k)
		{
			return std::move(target_[k]) / chi_->eval(chi_->sample_point(k));
		}
	};

	class PolynomialProgram
	{
		uint32_t N_;
		mp::real obj_const_{};
		algebra::Vector<mp::real> obj_;
		std::vector<algebra::Vector<mp::real>> equation_{};
		std::vector<mp::real> equation_targets_{};
		std::vector<uint32_t> leading_indices_{};  // pivots of the Gaussian elimination: y[leading_indices_[e]] is not free
		std::vector<uint32_t> free_indices_{};
		std::vector<std::optional<PolynomialInequality>> inequality_{};
		std::shared_ptr<b200::Engine> engine_{};   // where create_input runs; taken from the first device-resident inequality
		// objective after elimination: (constant, coefficients of the free variables)
		std::pair<mp::real, algebra::Vector<mp::real>> eliminated_objective();

	public:
		void _reset() &&
		{
			N_ = 0;
			std::move(obj_const_)._reset();
			std::move(obj_)._reset();
			std::vector<algebra::Vector<mp::real>>{}.swap(equation_);
			std::vector<mp::real>{}.swap(equation_targets_);
			std::vector<uint32_t>{}.swap(leading_indices_);
			std::vector<uint32_t>{}.swap(free_indices_);
			std::vector<std::optional<PolynomialInequality>>{}.swap(inequality_);
			engine_.reset();
		}
		explicit PolynomialProgram(uint32_t num_of_vars) : N_(num_of_vars), obj_(num_of_vars)
		{
			for (uint32_t i = 0; i < N_; ++i) free_indices_.push_back(i);
		}
		[[nodiscard]] uint32_t num_of_variables() const { return N_; }
		[[nodiscard]] uint32_t num_of_inequalities() const { return uint32_t(inequality_.size()); }
		[[nodiscard]] uint32_t _total_memory() const
		{
			uint32_t sum = 1 + obj_.size();
			for (const auto& x : inequality_)
				if (x) sum += x->_total_memory();
			for (const auto& x : equation_) sum += x.size() + 1;
			return sum;
		}
		[[nodiscard]] mp::real& objective_constant() { return obj_const_; }
		[[nodiscard]] const mp::real& objective_constant() const { return obj_const_; }
		[[nodiscard]] const algebra::Vector<mp::real>& objectives() const { return obj_; }
		void objectives(algebra::Vector<mp::real>&& obj) &
		{
			assert(obj.size() == N_);
			obj_ = std::move(obj);
		}
		void _use_engine(std::shared_ptr<b200::Engine> e) & { engine_ = std::move(e); }
		// y (all N variables) from z (the free ones) through the equations
		algebra::Vector<mp::real> recover(const algebra::Vector<mp::real>& z);
		// add sum_n y[n] vec[n] = target; equations must be independent; the order of calls fixes the output
		void add_equation(algebra::Vector<mp::real>&& vec, mp::real&& target) &;
		void add_inequality(std::optional<PolynomialInequality>&& ineq) &
		{
			assert(ineq->num_of_variables() == N_);
			inequality_.push_back(std::move(ineq));
		}
		SDPBInput create_input(uint32_t parallel = 1, const std::unique_ptr<_event_base>& event = {}) &&;
		XMLInput create_xml(uint32_t parallel = 1, const std::unique_ptr<_event_base>& event = {}) &&;
	};
}  // namespace qboot

#endif  // QBOOT_B200_POLYNOMIAL_PROGRAM_HPP_
