// qboot_b200 façade -- one term of a crossing equation: an exchanged operator between four externals
// (call surface of include/qboot/block.hpp).  From the external dimensions the engine needs
//   S = (d34 - d12) / 2,  P = -d12 d34 / 2  (keys of a block table)   and   d23h = (d2 + d3) / 2  (the power of v),
// each computed here with the reference's operations (block.hpp:15-24) so the table keys are the same bits.
#ifndef QBOOT_B200_BLOCK_HPP_
#define QBOOT_B200_BLOCK_HPP_

#include <cassert>
#include <string>
#include <type_traits>
#include <utility>
#include <variant>

#include "qboot/algebra/complex_function.hpp"
#include "qboot/mp/real.hpp"
#include "qboot/primary_op.hpp"

namespace qboot
{
	inline mp::real _delta_S(const mp::real& d12, const mp::real& d34) { return (d34 - d12) / 2; }
	inline mp::real _delta_S(const mp::real& d1, const mp::real& d2, const mp::real& d3, const mp::real& d4) { return _delta_S(d1 - d2, d3 - d4); }
	inline mp::real _delta_P(const mp::real& d12, const mp::real& d34) { return d12 * d34 / (-2); }
	inline mp::real _delta_P(const mp::real& d1, const mp::real& d2, const mp::real& d3, const mp::real& d4) { return _delta_P(d1 - d2, d3 - d4); }

	// what every block carries besides its operator
	class _block_data
	{
	protected:
		mp::real d12_, d34_, d23h_, S_{}, P_{};
		algebra::FunctionSymmetry sym_;  // Odd: F (F_-), Even: H (F_+)
		_block_data(const mp::real& d12, const mp::real& d34, const mp::real& d23h, algebra::FunctionSymmetry sym)
		    : d12_(d12), d34_(d34), d23h_(d23h), S_(_delta_S(d12, d34)), P_(_delta_P(d12, d34)), sym_(sym)
		{
			assert(sym == algebra::FunctionSymmetry::Even || sym == algebra::FunctionSymmetry::Odd);
		}
		void clear()
		{
			d12_ = d34_ = d23h_ = S_ = P_ = mp::real{};
			sym_ = algebra::FunctionSymmetry::Mixed;
		}
		[[nodiscard]] std::string describe(const std::string& op) const;

	public:
		[[nodiscard]] const mp::real& delta_half() const { return d23h_; }
		[[nodiscard]] const mp::real& S() const { return S_; }
		[[nodiscard]] const mp::real& P() const { return P_; }
		[[nodiscard]] bool include_odd() const { return d12_ != 0 && d34_ != 0; }
		[[nodiscard]] algebra::FunctionSymmetry symmetry() const { return sym_; }
	};
	// F^{d1 d2, d3 d4}_{-+, op}
	template <class Operator>
	class ConformalBlock : public _block_data
	{
		Operator op_;

	public:
		void _reset() &&
		{
			std::move(op_)._reset();
			clear();
		}
		ConformalBlock(const Operator& op, const mp::real& d12, const mp::real& d34, const mp::real& d23h,
		               algebra::FunctionSymmetry sym = algebra::FunctionSymmetry::Odd)
		    : _block_data(d12, d34, d23h, sym), op_(op)
		{
		}
		ConformalBlock(const Operator& op, const mp::real& d1, const mp::real& d2, const mp::real& d3, const mp::real& d4,
		               algebra::FunctionSymmetry sym = algebra::FunctionSymmetry::Odd)
		    : ConformalBlock(op, d1 - d2, d3 - d4, (d2 + d3) / 2, sym)
		{
		}
		ConformalBlock(const Operator& op, const PrimaryOperator& o1, const PrimaryOperator& o2, const PrimaryOperator& o3,
		               const PrimaryOperator& o4, algebra::FunctionSymmetry sym = algebra::FunctionSymmetry::Odd)
		    : ConformalBlock(op, o1.delta(), o2.delta(), o3.delta(), o4.delta(), sym)
		{
		}
		template <class O = Operator, class = std::enable_if_t<std::is_same_v<O, PrimaryOperator>>>
		[[nodiscard]] const mp::real& delta() const
		{
			return op_.delta();
		}
		[[nodiscard]] uint32_t spin() const { return op_.spin(); }
		[[nodiscard]] const mp::rational& epsilon() const { return op_.epsilon(); }
		[[nodiscard]] const Operator& get_op() const { return op_; }
		template <class O = Operator, class = std::enable_if_t<!std::is_same_v<O, PrimaryOperator>>>
		[[nodiscard]] PrimaryOperator get_op(const mp::real& delta) const
		{
			return op_.fix_delta(delta);
		}
		template <class O = Operator, class = std::enable_if_t<!std::is_same_v<O, PrimaryOperator>>>
		[[nodiscard]] ConformalBlock<PrimaryOperator> fix_delta(const mp::real& delta) const
		{
			return ConformalBlock<PrimaryOperator>(op_.fix_delta(delta), d12_, d34_, d23h_, sym_);
		}
		[[nodiscard]] std::string str() const { return describe("op=" + op_.str() + ", "); }
	};
	// a block whose exchanged operator is still open (a continuous sector supplies it)
	class GeneralConformalBlock : public _block_data
	{
	public:
		void _reset() && { clear(); }
		GeneralConformalBlock(const mp::real& d12, const mp::real& d34, const mp::real& d23h,
		                      algebra::FunctionSymmetry sym = algebra::FunctionSymmetry::Odd)
		    : _block_data(d12, d34, d23h, sym)
		{
		}
		GeneralConformalBlock(const mp::real& d1, const mp::real& d2, const mp::real& d3, const mp::real& d4,
		                      algebra::FunctionSymmetry sym = algebra::FunctionSymmetry::Odd)
		    : GeneralConformalBlock(d1 - d2, d3 - d4, (d2 + d3) / 2, sym)
		{
		}
		GeneralConformalBlock(const PrimaryOperator& o1, const PrimaryOperator& o2, const PrimaryOperator& o3,
		                      const PrimaryOperator& o4, algebra::FunctionSymmetry sym = algebra::FunctionSymmetry::Odd)
		    : GeneralConformalBlock(o1.delta(), o2.delta(), o3.delta(), o4.delta(), sym)
		{
		}
		template <class Operator>
		[[nodiscard]] ConformalBlock<Operator> fix_op(const Operator& op) const
		{
			return ConformalBlock<Operator>(op, d12_, d34_, d23h_, sym_);
		}
		[[nodiscard]] std::string str() const { return describe(""); }
	};
	using Block = std::variant<ConformalBlock<PrimaryOperator>, GeneralConformalBlock>;
}  // namespace qboot

#endif  // QBOOT_B200_BLOCK_HPP_
