// qboot_b200 façade -- qboot::Context on the B200 engine (call surface of include/qboot/context.hpp).
//
// In the reference a Context owns two memo tables, each with its own pool of worker threads, and every block table is
// computed by one MPFR thread on demand (context.hpp:29-56, src/context.cpp:47-61).  Here a Context owns one engine
// handle (qb_create / qb_context_init of include/qboot_b200.h) and the tables are computed on the GPU.  Single calls
// (gBlock, evaluate, F_block ...) still work and are memoised by key, but the intended path is the batched one:
// BootstrapEquation::convert collects every (Delta, spin, S, P) key of a whole program, removes duplicates -- the role of
// the reference's _memoized -- and evaluates them in one launch sequence (qboot_b200/host/facade_program.cpp).
//
// The handle is created on first use, so model descriptions, scale factors and sample points can be built without a
// GPU; the first request for a table on a machine without a CUDA device throws (there is no CPU path for tables).
#ifndef QBOOT_B200_CONTEXT_HPP_
#define QBOOT_B200_CONTEXT_HPP_

#include <cstdint>
#include <memory>
#include <string>

#include "qboot/algebra/complex_function.hpp"
#include "qboot/algebra/real_function.hpp"
#include "qboot/block.hpp"
#include "qboot/mp/rational.hpp"
#include "qboot/mp/real.hpp"
#include "qboot/primary_op.hpp"

namespace qboot
{
	namespace b200
	{
		class Engine;  // RAII owner of one qb_handle (qboot_b200/host/engine_link.hpp)
	}
	class Context
	{
		uint32_t n_Max_, lambda_, parallel_;
		mp::rational dim_, epsilon_;
		mp::real rho_;
		struct Memo;
		std::unique_ptr<Memo> memo_;  // engine handle, rho -> z matrix, memoised tables

	public:
		Context(uint32_t n_Max, uint32_t lambda, const mp::rational& dim, uint32_t p = 1);
		Context(Context&&) noexcept = delete;
		Context& operator=(Context&&) noexcept = delete;
		Context(const Context&) = delete;
		Context& operator=(const Context&) = delete;
		~Context();
		// cutoff of the rho-series, derivative order, spacetime dimension, epsilon = (dim - 2) / 2, rho* = 3 - 2 sqrt 2
		[[nodiscard]] uint32_t n_Max() const { return n_Max_; }
		[[nodiscard]] uint32_t lambda() const { return lambda_; }
		[[nodiscard]] const mp::rational& dimension() const { return dim_; }
		[[nodiscard]] uint32_t parallel() const { return parallel_; }
		[[nodiscard]] const mp::rational& epsilon() const { return epsilon_; }
		[[nodiscard]] const mp::real& rho() const { return rho_; }
		[[nodiscard]] const algebra::RealConverter& rho_to_z() const;
		[[nodiscard]] std::string str() const;
		// number of reals held in the memo tables (+ the conversion matrix + 1), as the reference counts them
		[[nodiscard]] uint32_t _total_memory() const;
		[[nodiscard]] mp::rational unitarity_bound(uint32_t spin) const { return qboot::unitarity_bound(epsilon_, spin); }
		// the engine behind this Context (created on first call; throws std::runtime_error without a CUDA device)
		[[nodiscard]] std::shared_ptr<b200::Engine> _engine() const;

		// ((1 - z)(1 - zbar))^d as a table, memoised per d
		[[nodiscard]] const algebra::ComplexFunction<mp::real>& v_to_d(const mp::real& d) const;
		// the block-derivative table of (op; S, P), memoised per key
		[[nodiscard]] const algebra::ComplexFunction<mp::real>& gBlock(const PrimaryOperator& op, const mp::real& S, const mp::real& P) const;
		[[nodiscard]] const algebra::ComplexFunction<mp::real>& gBlock(const PrimaryOperator& op, const mp::real& d1, const mp::real& d2,
		                                                               const mp::real& d3, const mp::real& d4) const
		{
			return gBlock(op, _delta_S(d1, d2, d3, d4), _delta_P(d1, d2, d3, d4));
		}
		// 2 proj_sym(v^d g): sym Odd = F-type, Even = H-type
		[[nodiscard]] algebra::ComplexFunction<mp::real> F_block(const mp::real& d, const algebra::ComplexFunction<mp::real>& gBlock,
		                                                         algebra::FunctionSymmetry sym) const;
		[[nodiscard]] algebra::ComplexFunction<mp::real> evaluate(const ConformalBlock<PrimaryOperator>& block) const;
		[[nodiscard]] algebra::ComplexFunction<mp::real> evaluate(const ConformalBlock<GeneralPrimaryOperator>& block, const mp::real& delta) const;
		[[nodiscard]] algebra::ComplexFunction<mp::real> F_block(const PrimaryOperator& op, const mp::real& d1, const mp::real& d2,
		                                                         const mp::real& d3, const mp::real& d4, algebra::FunctionSymmetry sym) const
		{
			return evaluate(ConformalBlock<PrimaryOperator>(op, d1, d2, d3, d4, sym));
		}
		[[nodiscard]] algebra::ComplexFunction<mp::real> F_block(const mp::real& d2, const mp::real& d3, const algebra::ComplexFunction<mp::real>& g) const
		{
			return F_block((d2 + d3) / 2, g);
		}
		[[nodiscard]] algebra::ComplexFunction<mp::real> F_block(const mp::real& d, const algebra::ComplexFunction<mp::real>& g) const
		{
			return F_block(d, g, algebra::FunctionSymmetry::Odd);
		}
		[[nodiscard]] algebra::ComplexFunction<mp::real> F_block(const PrimaryOperator& op, const mp::real& d1, const mp::real& d2,
		                                                         const mp::real& d3, const mp::real& d4) const
		{
			return F_block(op, d1, d2, d3, d4, algebra::FunctionSymmetry::Odd);
		}
		[[nodiscard]] algebra::ComplexFunction<mp::real> H_block(const mp::real& d2, const mp::real& d3, const algebra::ComplexFunction<mp::real>& g) const
		{
			return H_block((d2 + d3) / 2, g);
		}
		[[nodiscard]] algebra::ComplexFunction<mp::real> H_block(const mp::real& d, const algebra::ComplexFunction<mp::real>& g) const
		{
			return F_block(d, g, algebra::FunctionSymmetry::Even);
		}
		[[nodiscard]] algebra::ComplexFunction<mp::real> H_block(const PrimaryOperator& op, const mp::real& d1, const mp::real& d2,
		                                                         const mp::real& d3, const mp::real& d4) const
		{
			return F_block(op, d1, d2, d3, d4, algebra::FunctionSymmetry::Even);
		}
	};
	// the free functions of include/qboot/hor_recursion.hpp:19-30 that user code may call directly
	algebra::ComplexFunction<mp::real> gBlock(const PrimaryOperator& op, const mp::real& S, const mp::real& P, const Context& context);
	algebra::ComplexFunction<mp::real> gBlock(const PrimaryOperator& op, const mp::real& d1, const mp::real& d2, const mp::real& d3,
	                                          const mp::real& d4, const Context& context);
}  // namespace qboot

#endif  // QBOOT_B200_CONTEXT_HPP_
