// qboot_b200 façade -- the description of a bootstrap problem and its conversion into a polynomial matrix program
// (call surface of include/qboot/bootstrap_equation.hpp).
//
// Sector / Equation / BootstrapEquation describe the crossing equations exactly as user code written for the reference
// does.  `convert` is where the engines differ.  The reference walks (sector, operator) -> equation -> sample point ->
// term on `parallel` threads and asks a memoised Context for one block table at a time (src/bootstrap_equation.cpp:
// 423-461,555-593).  Here the same walk only PLANS: it collects every (Delta_k, spin, S, P) key of the whole program,
// removes duplicates (the memo's job), hands the list to the GPU in one batch (qb_gblock_plan / qb_gblock_run) and then
// lets one kernel form all F/H products and per-equation sums (qb_pmp_assemble).  The resulting PolynomialProgram
// refers to device-resident constraint blocks.  Details in qboot_b200/host/facade_program.cpp.
#ifndef QBOOT_B200_BOOTSTRAP_EQUATION_HPP_
#define QBOOT_B200_BOOTSTRAP_EQUATION_HPP_

#include <algorithm>
#include <array>
#include <cassert>
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <optional>
#include <string>
#include <string_view>
#include <tuple>
#include <utility>
#include <variant>
#include <vector>

#include "qboot/algebra/complex_function.hpp"
#include "qboot/algebra/matrix.hpp"
#include "qboot/block.hpp"
#include "qboot/conformal_scale.hpp"
#include "qboot/context.hpp"
#include "qboot/mp/rational.hpp"
#include "qboot/mp/real.hpp"
#include "qboot/my_filesystem.hpp"
#include "qboot/polynomial_program.hpp"
#include "qboot/primary_op.hpp"
#include "qboot/task_queue.hpp"

namespace qboot
{
	// spin -> number of poles kept: numax + min(numax, spin) / 2
	class _pole_selector
	{
		uint32_t numax_;

	public:
		explicit _pole_selector(uint32_t numax) : numax_(numax) {}
		uint32_t operator()(uint32_t spin) const { return numax_ + std::min(numax_, spin) / 2; }
	};
	// One term of a crossing equation inside a sector: weight * ope[row] * ope[col] * block.
	class Entry
	{
		struct Position
		{
			uint32_t row = 0, col = 0;
		};
		Position at_{};
		mp::real weight_;
		Block what_;

	public:
		// (row, col, weight, block); position (0, 0) and weight 1 when left out
		Entry(uint32_t r, uint32_t c, const mp::real& coeff, const Block& block) : at_{r, c}, weight_(coeff), what_(block) {}
		Entry(uint32_t r, uint32_t c, const Block& block) : at_{r, c}, weight_(1), what_(block) {}
		Entry(const mp::real& coeff, const Block& block) : weight_(coeff), what_(block) {}
		explicit Entry(const Block& block) : weight_(1), what_(block) {}
		void _reset() &&
		{
			at_ = {};
			std::move(weight_)._reset();
			std::visit([](auto&& alt) { std::move(alt)._reset(); }, std::move(what_));
		}
		[[nodiscard]] uint32_t row() const { return at_.row; }
		[[nodiscard]] uint32_t column() const { return at_.col; }
		[[nodiscard]] const mp::real& coeff() const { return weight_; }
		[[nodiscard]] const Block& block() const { return what_; }
	};
	enum class SectorType : bool
	{
		Continuous = false,
		Discrete = true
	};
	// A set of operators sharing one size x size matrix of OPE coefficients.  A continuous sector lists (spin, range)
	// requests; they become GeneralPrimaryOperators when the BootstrapEquation that owns the sector knows epsilon and the
	// rule for the number of poles.  Bodies that are not one-liners: qboot_b200/host/facade_program.cpp.
	class Sector
	{
		friend class BootstrapEquation;
		struct Request  // Delta in [from, to); absent from: the unitarity bound, absent to: unbounded
		{
			uint32_t spin;
			std::optional<mp::real> from{}, to{};
		};
		std::string name_;
		uint32_t sz_;
		SectorType type_;
		std::vector<Request> requested_{};
		std::vector<GeneralPrimaryOperator> ops_{};
		std::optional<algebra::Vector<mp::real>> ope_{};
		void request(Request&& r) &;
		void set_operators(const mp::rational& epsilon, const std::function<uint32_t(uint32_t)>& num_poles) &;

	public:
		Sector(std::string_view name, uint32_t size, SectorType type = SectorType::Discrete);
		// a discrete sector whose OPE coefficients are known: contributes the number ope^t M ope
		Sector(std::string_view name, uint32_t size, algebra::Vector<mp::real>&& ope);
		~Sector() = default;
		Sector(Sector&&) = default;
		Sector& operator=(Sector&&) = default;
		Sector(const Sector& s);  // deep: the OPE vector is move-only and gets cloned
		Sector& operator=(const Sector& s) { return this == &s ? *this : (*this = Sector(s)); }
		void _reset() &&;
		[[nodiscard]] const std::string& name() const { return name_; }
		[[nodiscard]] uint32_t size() const { return sz_; }
		[[nodiscard]] SectorType type() const { return type_; }
		[[nodiscard]] bool is_matrix() const { return sz_ > 1 && !ope_.has_value(); }
		[[nodiscard]] const std::optional<algebra::Vector<mp::real>>& ope() const { return ope_; }
		[[nodiscard]] const std::vector<GeneralPrimaryOperator>& _operators() const { return ops_; }
		// continuous sectors only: Delta >= unitarity bound; Delta >= lb; lb <= Delta < ub (an infinite ub means none)
		void add_op(uint32_t spin) & { request(Request{spin}); }
		void add_op(uint32_t spin, const mp::real& lb) & { request(Request{spin, lb}); }
		void add_op(uint32_t spin, const mp::real& lb, const mp::real& ub) &
		{
			request(ub.isinf() ? Request{spin, lb} : Request{spin, lb, ub});
		}
	};
	class BootstrapEquation;
	// what to do with the functional: which sector normalises it, which is optimised (user-extensible)
	class SDPMode
	{
	public:
		SDPMode() = default;
		SDPMode(const SDPMode&) = default;
		SDPMode& operator=(const SDPMode&) = default;
		SDPMode(SDPMode&&) noexcept = default;
		SDPMode& operator=(SDPMode&&) noexcept = default;
		virtual ~SDPMode();
		virtual PolynomialProgram _prepare(uint32_t N, const BootstrapEquation& boot) const = 0;
		[[nodiscard]] virtual std::function<bool(const std::string&)> _get_filter() const = 0;
	};
	// alpha(norm) = 1 and alpha(sec) >= 0 for every other sector
	class _find_contradiction : public SDPMode
	{
		std::string norm_;

	public:
		_find_contradiction(std::string_view norm) : norm_(norm) {}  // NOLINT
		~_find_contradiction() override;
		PolynomialProgram _prepare(uint32_t N, const BootstrapEquation& boot) const override;
		[[nodiscard]] std::function<bool(const std::string&)> _get_filter() const override
		{
			return [n = norm_](const std::string& s) { return s != n; };
		}
	};
	inline std::unique_ptr<SDPMode> FindContradiction(std::string_view norm) { return std::make_unique<_find_contradiction>(norm); }
	// maximise alpha(norm) subject to alpha(target) = +-1 (+1: upper bound on the OPE coefficient, -1: lower bound) and
	// alpha(sec) >= 0 for every other sector
	class _extremal_OPE : public SDPMode
	{
		bool maximize_;
		std::string target_, norm_;

	public:
		_extremal_OPE(bool maximize, std::string_view target, std::string_view norm) : maximize_(maximize), target_(target), norm_(norm) {}
		~_extremal_OPE() override;
		PolynomialProgram _prepare(uint32_t N, const BootstrapEquation& boot) const override;
		[[nodiscard]] std::function<bool(const std::string&)> _get_filter() const override
		{
			return [t = target_, n = norm_](const std::string& s) { return s != t && s != n; };
		}
	};
	inline std::unique_ptr<SDPMode> ExtremalOPE(bool maximize, std::string_view target, std::string_view norm)
	{
		return std::make_unique<_extremal_OPE>(maximize, target, norm);
	}
	// y.txt of an SDPB run -> the raw functional
	algebra::Vector<mp::real> read_raw_functional(const fs::path& y_txt);
	class Equation;
	// the Context must outlive the BootstrapEquation
	class BootstrapEquation
	{
		const Context& cont_;
		std::vector<Sector> sectors_{};
		std::map<std::string, uint32_t, std::less<>> sector_id_{};
		std::vector<Equation> eqs_{};
		uint32_t N_ = 0;  // sum of the equations' dimensions
		struct Plan;      // facade_program.cpp
		friend struct Plan;
		// values of discrete sectors computed ahead by convert() so that SDPMode::_prepare finds them ready
		mutable std::map<uint32_t, algebra::Vector<algebra::Matrix<mp::real>>> disc_cache_{};
		[[nodiscard]] mp::real take_element(uint32_t id, algebra::Matrix<mp::real>&& m) const
		{
			const auto& ope = sector(id).ope();
			if (ope) return m.inner_product(ope.value());
			return std::move(m.at(0, 0));
		}
		[[nodiscard]] bool sector_includes_odd(uint32_t id) const;

	public:
		void _reset() &&
		{
			sectors_.clear();
			sector_id_.clear();
			eqs_.clear();
			disc_cache_.clear();
			N_ = 0;
		}
		BootstrapEquation(const Context& cont, const std::vector<Sector>& sectors, uint32_t numax)
		    : BootstrapEquation(cont, std::vector<Sector>(sectors), _pole_selector(numax))
		{
		}
		BootstrapEquation(const Context& cont, std::vector<Sector>&& sectors, uint32_t numax)
		    : BootstrapEquation(cont, std::move(sectors), _pole_selector(numax))
		{
		}
		BootstrapEquation(const Context& cont, const std::vector<Sector>& sectors, const std::function<uint32_t(uint32_t)>& num_poles)
		    : BootstrapEquation(cont, std::vector<Sector>(sectors), num_poles)
		{
		}
		BootstrapEquation(const Context& cont, std::vector<Sector>&& sectors, const std::function<uint32_t(uint32_t)>& num_poles)
		    : cont_(cont), sectors_(std::move(sectors))
		{
			for (uint32_t id = 0; id < sectors_.size(); ++id)
			{
				sector_id_[sectors_[id].name()] = id;
				if (sectors_[id].type() == SectorType::Continuous) sectors_[id].set_operators(cont.epsilon(), num_poles);
			}
		}
		void add_equation(const Equation& eq) &;
		void add_equation(Equation&& eq) &;
		void finish() &;  // call after the last add_equation
		[[nodiscard]] const Context& context() const { return cont_; }
		[[nodiscard]] uint32_t lambda() const { return cont_.lambda(); }
		[[nodiscard]] uint32_t num_sectors() const { return uint32_t(sectors_.size()); }
		[[nodiscard]] uint32_t num_variables() const { return N_; }
		[[nodiscard]] uint32_t get_id(std::string_view sector_name) const { return sector_id_.find(sector_name)->second; }
		[[nodiscard]] const Sector& sector(std::string_view sector_name) const { return sectors_[get_id(sector_name)]; }
		[[nodiscard]] const Sector& sector(uint32_t id) const { return sectors_[id]; }
		// N matrices (sz x sz) of a discrete sector, and their contraction with the sector's OPE vector
		[[nodiscard]] algebra::Vector<algebra::Matrix<mp::real>> make_disc_mat(uint32_t id) const;
		[[nodiscard]] algebra::Vector<algebra::Matrix<mp::real>> make_disc_mat(std::string_view sector_name) const
		{
			return make_disc_mat(get_id(sector_name));
		}
		[[nodiscard]] algebra::Vector<mp::real> make_disc_mat_v(uint32_t id) const
		{
			assert(!sector(id).is_matrix() && sector(id).type() == SectorType::Discrete);
			auto m = make_disc_mat(id);
			algebra::Vector<mp::real> v(N_);
			for (uint32_t i = 0; i < N_; ++i) v[i] = take_element(id, std::move(m[i]));
			return v;
		}
		[[nodiscard]] algebra::Vector<mp::real> make_disc_mat_v(std::string_view sector_name) const { return make_disc_mat_v(get_id(sector_name)); }
		// the scale factor and the N x (D + 1) sample matrices of one operator of a continuous sector
		[[nodiscard]] std::unique_ptr<ConformalScale> common_scale(uint32_t id, const GeneralPrimaryOperator& op) const;
		[[nodiscard]] algebra::Vector<algebra::Vector<algebra::Matrix<mp::real>>> make_cont_mat(uint32_t id, const GeneralPrimaryOperator& op,
		                                                                                      const std::unique_ptr<ConformalScale>& ag) const;
		algebra::Matrix<mp::real> apply_functional(const algebra::Vector<mp::real>& func, std::string_view disc_sector) const;
		[[nodiscard]] PolynomialProgram convert(const std::unique_ptr<SDPMode>& mode, uint32_t parallel = 1,
		                                        const std::unique_ptr<_event_base>& event = {}) const;
		[[nodiscard]] algebra::Vector<mp::real> recover(const std::unique_ptr<SDPMode>& mode, const algebra::Vector<mp::real>& raw_func) const;
		// Extremal functional method (src/bootstrap_equation.cpp:463-542).  Post-processing of an SDPB solution by exact
		// rational root isolation: outside the scope of this engine.  Declared so that drivers written
		// for the reference compile; calling it throws std::logic_error.
		std::map<std::string, std::vector<PrimaryOperator>, std::less<>> get_spectrum(const algebra::Vector<mp::real>& recovered_func,
		                                                                              uint32_t parallel = 1,
		                                                                              const std::unique_ptr<_event_base>& event = {}) const;
	};
	using Externals = std::array<PrimaryOperator, 4>;
	// one crossing equation: terms per sector, all of one symmetry (Odd = F-type, Even = H-type)
	class Equation
	{
		const BootstrapEquation& boot_;
		algebra::FunctionSymmetry sym_;
		uint32_t dim_;
		std::vector<std::vector<Entry>> terms_{};  // terms_[sector id]

	public:
		void _reset() &&
		{
			sym_ = algebra::FunctionSymmetry::Mixed;
			dim_ = 0;
			terms_.clear();
		}
		Equation(const BootstrapEquation& boot, algebra::FunctionSymmetry sym)
		    : boot_(boot), sym_(sym), dim_(algebra::function_dimension(boot.lambda(), sym)), terms_(boot.num_sectors())
		{
		}
		[[nodiscard]] algebra::FunctionSymmetry symmetry() const { return sym_; }
		[[nodiscard]] uint32_t dimension() const { return dim_; }
		const std::vector<Entry>& operator[](uint32_t id) const { return terms_[id]; }
		// a fixed operator o exchanged between the externals os, at position (r, c) of the sector's matrix
		void add(std::string_view sec, uint32_t r, uint32_t c, const mp::real& coeff, const PrimaryOperator& o, const Externals& os)
		{
			const auto id = boot_.get_id(sec);
			assert(r < boot_.sector(id).size() && c < boot_.sector(id).size());
			terms_[id].emplace_back(r, c, coeff, ConformalBlock<PrimaryOperator>(o, os[0], os[1], os[2], os[3], sym_));
		}
		void add(std::string_view sec, const mp::real& coeff, const PrimaryOperator& o, const Externals& os) { add(sec, 0, 0, coeff, o, os); }
		void add(std::string_view sec, uint32_t r, uint32_t c, const PrimaryOperator& o, const Externals& os) { add(sec, r, c, mp::real(1), o, os); }
		void add(std::string_view sec, const PrimaryOperator& o, const Externals& os) { add(sec, 0, 0, o, os); }
		// the operators of a continuous sector exchanged between the externals os
		void add(std::string_view sec, uint32_t r, uint32_t c, const mp::real& coeff, const Externals& os)
		{
			const auto id = boot_.get_id(sec);
			assert(r < boot_.sector(id).size() && c < boot_.sector(id).size());
			terms_[id].emplace_back(r, c, coeff, GeneralConformalBlock(os[0], os[1], os[2], os[3], sym_));
		}
		void add(std::string_view sec, const mp::real& coeff, const Externals& os) { add(sec, 0, 0, coeff, os); }
		void add(std::string_view sec, uint32_t r, uint32_t c, const Externals& os) { add(sec, r, c, mp::real(1), os); }
		void add(std::string_view sec, const Externals& os) { add(sec, 0, 0, os); }
	};
	inline void BootstrapEquation::add_equation(const Equation& eq) &
	{
		assert(N_ == 0);
		eqs_.push_back(eq);
	}
	inline void BootstrapEquation::add_equation(Equation&& eq) &
	{
		assert(N_ == 0);
		eqs_.push_back(std::move(eq));
	}
}  // namespace qboot

#endif  // QBOOT_B200_BOOTSTRAP_EQUATION_HPP_
