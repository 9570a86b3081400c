// qboot_b200 host library -- qboot::Context and the engine link (include/qboot/context.hpp).
//
// Replaces src/context.cpp:47-61 (Context::Context) and the memoised callables of include/qboot/context.hpp:29-56: the
// Context owns one engine handle; single-table requests are memoised by key and served by one-element batches, whole
// programs go through the planner in facade_program.cpp.
#include <cstdlib>
#include <cstring>
#include <map>
#include <sstream>
#include <vector>

#include "engine_link.hpp"
#include <mutex>

#include "qboot/context.hpp"
#include "qboot/primary_op.hpp"
#include "qboot/task_queue.hpp"

namespace qboot
{
	using algebra::ComplexFunction, algebra::FunctionSymmetry, algebra::Matrix, algebra::Vector;
	using mp::rational, mp::real;

	_event_base::~_event_base() = default;

	namespace b200
	{
		static int pick_device()
		{
			// one process per GPU: QB_DEVICE wins, else the launcher's LOCAL_RANK, else device 0
			for (const char* name : {"QB_DEVICE", "LOCAL_RANK"})
				if (const char* v = std::getenv(name)) return std::atoi(v);
			return 0;
		}
		namespace
		{
			// One spare engine handle per (device, precision) outlives the Context that used it: a scan builds one Context per
			// grid point, and creating a handle means allocating (and, on destruction, freeing) its device arenas -- hundreds
			// of megabytes, occasionally half a second.  qb_context_init resets everything a previous owner left behind.
			// The pool is never destroyed (no CUDA calls during static destruction).
			struct HandlePool
			{
				std::mutex mu;
				std::map<std::pair<int, int>, qb_handle*> spare;
			};
			HandlePool& handle_pool()
			{
				static HandlePool* p = new HandlePool;
				return *p;
			}
		}  // namespace
		Engine::Engine(int prec_bits, uint32_t n_max_, uint32_t lambda_, const rational& epsilon)
		    : prec(prec_bits), device(pick_device()), n_max(n_max_), lambda(lambda_)
		{
			qb::host::check_prec(prec);
			if (epsilon.den() > int64_t(UINT32_MAX)) throw std::overflow_error("qboot_b200: epsilon denominator exceeds 32 bits");
			{
				HandlePool& pool = handle_pool();
				std::lock_guard<std::mutex> g(pool.mu);
				auto it = pool.spare.find({device, prec});
				if (it != pool.spare.end())
				{
					h = it->second;
					pool.spare.erase(it);
				}
			}
			int rc = h ? QB_OK : qb_create(device, prec, &h);
			if (rc != QB_OK)
				throw std::runtime_error(std::string("qboot_b200: cannot create the GPU engine: ") + qb_strerror(rc) +
				                         " (block tables are computed on a CUDA device; there is no CPU path)");
			rc = qb_context_init(h, n_max, lambda, epsilon.num(), uint32_t(epsilon.den()));
			if (rc != QB_OK)
			{
				std::string msg = std::string("qboot_b200: qb_context_init failed: ") + qb_strerror(rc) + " " + qb_last_cuda_error(h);
				qb_destroy(h);
				h = nullptr;
				throw std::runtime_error(msg);
			}
			words = qb_words(h);
			dim_mixed = qb_table_dim(h);
			dim_sym[0] = qb_block_dim(h, 0);
			dim_sym[1] = qb_block_dim(h, 1);
		}
		Engine::~Engine()
		{
			if (!h) return;
			qb_handle* drop = h;
			{
				HandlePool& pool = handle_pool();
				std::lock_guard<std::mutex> g(pool.mu);
				auto& slot = pool.spare[{device, prec}];
				if (!slot)
				{
					slot = h;
					drop = nullptr;
				}
			}
			if (drop) qb_destroy(drop);
		}
		int Engine::check(int rc, const char* what, bool nonneg_ok) const
		{
			if (rc == QB_OK || (nonneg_ok && rc >= 0)) return rc;
			std::string msg = std::string("qboot_b200: ") + what + ": " + qb_strerror(rc);
			if (rc == QB_ERR_CUDA) msg += std::string(" [") + qb_last_cuda_error(h) + "]";
			throw std::runtime_error(msg);
		}
		Pinned::Pinned(size_t bytes) : p(qb_pinned_alloc(bytes))
		{
			if (!p && bytes) throw std::bad_alloc();
		}
		Pinned::~Pinned() { qb_pinned_free(p); }
		PackedProgram::~PackedProgram() { pinned_release(text, text_cap); }
		namespace
		{
			std::mutex pinned_mu;
			char* pinned_spare = nullptr;
			size_t pinned_spare_bytes = 0;
		}  // namespace
		char* pinned_acquire(size_t bytes, size_t& capacity)
		{
			capacity = 0;
			if (bytes == 0) return nullptr;
			{
				std::lock_guard<std::mutex> g(pinned_mu);
				if (pinned_spare && pinned_spare_bytes >= bytes)
				{
					char* p = pinned_spare;
					capacity = pinned_spare_bytes;
					pinned_spare = nullptr;
					pinned_spare_bytes = 0;
					return p;
				}
			}
			char* p = static_cast<char*>(qb_pinned_alloc(bytes));
			if (!p) throw std::bad_alloc();
			capacity = bytes;
			return p;
		}
		void pinned_release(char* p, size_t bytes)
		{
			if (!p) return;
			char* drop = p;
			{
				std::lock_guard<std::mutex> g(pinned_mu);
				if (bytes > pinned_spare_bytes)
				{
					drop = pinned_spare;
					pinned_spare = p;
					pinned_spare_bytes = bytes;
				}
			}
			qb_pinned_free(drop);
		}
		void format_numbers(Engine& e, int ndigits, const uint32_t* recs, size_t n, std::string& out, std::vector<size_t>* ends)
		{
			if (ends) ends->reserve(ends->size() + n);
			if (n == 0) return;
			const int slot = ndigits + 16;
			std::vector<char> slots(n * size_t(slot));
			std::vector<int32_t> len(n);
			out.reserve(out.size() + n * size_t(ndigits + 8));
			{
				std::lock_guard<std::recursive_mutex> g(e.mu);
				e.check(qb_dec_format(e.h, ndigits, int(n), recs, slots.data(), slot, len.data()), "qb_dec_format");
			}
			for (size_t i = 0; i < n; ++i)
			{
				if (len[i] >= 0)
					out.append(slots.data() + i * size_t(slot), size_t(len[i]));
				else
				{
					real r;
					std::memcpy(r._words(), recs + i * size_t(e.words), 4 * size_t(e.words));
					out += mp::_format(r, std::ios_base::fmtflags{}, ndigits);
					out += '\n';
				}
				if (ends) ends->push_back(out.size());
			}
		}
	}  // namespace b200

	// ---- Context ----------------------------------------------------------------------------------------------------------
	struct Context::Memo
	{
		std::mutex mu;
		std::shared_ptr<b200::Engine> engine;
		std::unique_ptr<algebra::RealConverter> rho_to_z;
		std::map<std::vector<uint32_t>, ComplexFunction<real>> tables, vpowers;
	};
	// src/context.cpp:47-61: epsilon = dim / 2 - 1, rho = 3 - sqrt(8); the rho -> z matrix is built by the engine
	Context::Context(uint32_t n_Max, uint32_t lambda, const rational& dim, uint32_t p)
	    : n_Max_(n_Max), lambda_(lambda), parallel_(p), dim_(dim), epsilon_(dim / 2 - 1), rho_(3 - mp::sqrt(8)), memo_(std::make_unique<Memo>())
	{
		qb::host::check_prec(mp::_prec());
	}
	Context::~Context() = default;
	std::shared_ptr<b200::Engine> Context::_engine() const
	{
		std::lock_guard<std::mutex> g(memo_->mu);
		if (!memo_->engine) memo_->engine = std::make_shared<b200::Engine>(mp::_prec(), n_Max_, lambda_, epsilon_);
		return memo_->engine;
	}
	const algebra::RealConverter& Context::rho_to_z() const
	{
		auto e = _engine();
		std::lock_guard<std::mutex> g(memo_->mu);
		if (!memo_->rho_to_z)
		{
			const size_t W = size_t(e->words), n1 = size_t(lambda_) + 1;
			std::vector<uint32_t> buf(W * n1 * n1);
			e->check(qb_context_get(e->h, nullptr, buf.data()), "qb_context_get");
			Matrix<real> m{uint32_t(n1), uint32_t(n1)};
			for (size_t i = 0; i < n1 * n1; ++i) std::memcpy(m.at(uint32_t(i / n1), uint32_t(i % n1))._words(), buf.data() + W * i, 4 * W);
			memo_->rho_to_z = std::make_unique<algebra::RealConverter>(std::move(m), lambda_);
		}
		return *memo_->rho_to_z;
	}
	std::string Context::str() const
	{
		std::ostringstream os;
		os << "Context(nMax=" << n_Max_ << ", lambda=" << lambda_ << ", dim=" << dim_ << ")";
		return os.str();
	}
	uint32_t Context::_total_memory() const
	{
		std::lock_guard<std::mutex> g(memo_->mu);
		const uint32_t dimM = algebra::function_dimension(lambda_);
		return 1 + uint32_t(memo_->tables.size() + memo_->vpowers.size()) * dimM + (lambda_ + 1) * (lambda_ + 1);
	}
	namespace
	{
		ComplexFunction<real> table_from_records(const uint32_t* recs, size_t W, uint32_t lambda, FunctionSymmetry sym)
		{
			const uint32_t n = algebra::function_dimension(lambda, sym);
			Vector<real> v(n);
			for (uint32_t i = 0; i < n; ++i) std::memcpy(v[i]._words(), recs + W * i, 4 * W);
			return ComplexFunction<real>(std::move(v), lambda, sym);
		}
		void append_record(std::vector<uint32_t>& key, const real& x, size_t W) { key.insert(key.end(), x._words(), x._words() + W); }
	}  // namespace
	const ComplexFunction<real>& Context::gBlock(const PrimaryOperator& op, const real& S, const real& P) const
	{
		auto e = _engine();
		const size_t W = size_t(e->words);
		std::vector<uint32_t> key;
		key.reserve(3 * W + 1);
		append_record(key, op.delta(), W);
		key.push_back(op.spin());
		append_record(key, S, W);
		append_record(key, P, W);
		{
			std::lock_guard<std::mutex> g(memo_->mu);
			auto it = memo_->tables.find(key);
			if (it != memo_->tables.end()) return it->second;
		}
		std::vector<uint32_t> out(W * size_t(e->dim_mixed));
		const uint32_t spin = op.spin();
		{
			std::lock_guard<std::recursive_mutex> g(e->mu);
			e->check(qb_gblock_batch(e->h, 1, &spin, op.delta()._words(), S._words(), P._words(), out.data()), "qb_gblock_batch");
		}
		std::lock_guard<std::mutex> g(memo_->mu);
		return memo_->tables.emplace(std::move(key), table_from_records(out.data(), W, lambda_, FunctionSymmetry::Mixed)).first->second;
	}
	const ComplexFunction<real>& Context::v_to_d(const real& d) const
	{
		auto e = _engine();
		const size_t W = size_t(e->words);
		std::vector<uint32_t> key(d._words(), d._words() + W);
		{
			std::lock_guard<std::mutex> g(memo_->mu);
			auto it = memo_->vpowers.find(key);
			if (it != memo_->vpowers.end()) return it->second;
		}
		std::vector<uint32_t> out(W * size_t(e->dim_mixed));
		{
			std::lock_guard<std::recursive_mutex> g(e->mu);
			e->check(qb_vtod_batch(e->h, 1, d._words(), out.data()), "qb_vtod_batch");
		}
		std::lock_guard<std::mutex> g(memo_->mu);
		return memo_->vpowers.emplace(std::move(key), table_from_records(out.data(), W, lambda_, FunctionSymmetry::Mixed)).first->second;
	}
	// include/qboot/context.hpp:123-128, on the device: 2 proj_sym(v^d g)
	ComplexFunction<real> Context::F_block(const real& d, const ComplexFunction<real>& g, FunctionSymmetry sym) const
	{
		if (sym == FunctionSymmetry::Mixed || g.symmetry() != FunctionSymmetry::Mixed || g.lambda() != lambda_)
			throw std::invalid_argument("qboot::Context::F_block: needs a Mixed table of this Context and sym = Even or Odd");
		auto e = _engine();
		const size_t W = size_t(e->words);
		std::vector<uint32_t> tab(W * size_t(e->dim_mixed)), out(W * size_t(e->dim_sym[int(sym)]));
		for (uint32_t i = 0; i < g.size(); ++i) std::memcpy(tab.data() + W * i, g._packed()[i]._words(), 4 * W);
		const int32_t s = int32_t(sym);
		{
			std::lock_guard<std::recursive_mutex> gd(e->mu);
			e->check(qb_fblock_tables(e->h, 1, tab.data(), d._words(), &s, out.data()), "qb_fblock_tables");
		}
		return table_from_records(out.data(), W, lambda_, sym);
	}
	ComplexFunction<real> Context::evaluate(const ConformalBlock<PrimaryOperator>& block) const
	{
		return F_block(block.delta_half(), gBlock(block.get_op(), block.S(), block.P()), block.symmetry());
	}
	ComplexFunction<real> Context::evaluate(const ConformalBlock<GeneralPrimaryOperator>& block, const real& delta) const
	{
		return F_block(block.delta_half(), gBlock(block.get_op(delta), block.S(), block.P()), block.symmetry());
	}
	ComplexFunction<real> gBlock(const PrimaryOperator& op, const real& S, const real& P, const Context& context)
	{
		return context.gBlock(op, S, P).clone();
	}
	ComplexFunction<real> gBlock(const PrimaryOperator& op, const real& d1, const real& d2, const real& d3, const real& d4, const Context& context)
	{
		return context.gBlock(op, d1, d2, d3, d4).clone();
	}

	// ---- value types with text output (src/primary_op.cpp, src/block.cpp) -----------------------------------------------
	PrimaryOperator::PrimaryOperator(const Context& c) : PrimaryOperator(c.epsilon()) {}
	PrimaryOperator::PrimaryOperator(uint32_t spin, const Context& c) : PrimaryOperator(spin, c.epsilon()) {}
	PrimaryOperator::PrimaryOperator(const real& delta, uint32_t spin, const Context& c) : PrimaryOperator(delta, spin, c.epsilon()) {}
	std::string PrimaryOperator::str() const
	{
		std::ostringstream os;
		os << "operator(spin=" << spin() << ", dim=" << 2 * eps_ + 2 << ", delta=" << delta() << ")";
		return os.str();
	}
	std::string GeneralPrimaryOperator::str() const
	{
		std::ostringstream os;
		os << "operator(spin=" << spin_ << ", dim=" << 2 * eps_ + 2 << ", delta in [" << lower_bound() << ", " << (is_finite() ? upper_bound().str() : "infty")
		   << "))";
		return os.str();
	}
	std::string _block_data::describe(const std::string& op) const
	{
		std::ostringstream os;
		os << "ConformalBlock(" << op << "d12=" << d12_ << ", d34=" << d34_ << ", d23h=" << d23h_ << ", type=" << (sym_ == FunctionSymmetry::Odd ? 'F' : 'H')
		   << ")";
		return os.str();
	}
}  // namespace qboot
