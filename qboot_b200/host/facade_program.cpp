// qboot_b200 host library -- from a bootstrap problem to SDPB input files on the B200 engine.
//
// Reference path: BootstrapEquation::convert -> make_cont_mat / make_disc_mat (src/bootstrap_equation.cpp:380-461,555-593)
// -> PolynomialProgram::create_input (src/polynomial_program.cpp:152-232) -> SDPBInput::write (src/sdpb_input.cpp:138-199).
// There every term of every equation at every sample point asks a memoised Context for one block table and multiplies
// it on a host thread.  Here the host only PLANS (which tables, which terms, in which order) and the device does all of
// the arithmetic in the reference's order:
//   1. plan      walk sectors / operators / equations / terms as make_cont_mat does; collect the keys (Delta_k, spin, S, P),
//                drop duplicates (the role of the reference's _memoized); scale factors are built on host threads meanwhile
//   2. tables    qb_gblock_plan + qb_gblock_run: every distinct table once, resident in HBM
//   3. assemble  qb_pmp_assemble: F/H products and per-equation sums for every (constraint, position, sample, component)
//   4. pack      qb_pmp_pack: create_input's negation, re-indexing and pivot elimination
//   5. text      qb_pmp_text: decimal digits of d_B / d_c; the host writes the bytes to the files
// Only the normalisation vectors (a few N-vectors), the bilinear bases and file I/O stay on the host.
#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <fstream>
#include <map>
#include <mutex>
#include <numeric>
#include <sstream>
#include <thread>
#include <unordered_map>

#include "engine_link.hpp"
#include "qboot/bootstrap_equation.hpp"
#include "qboot/polynomial_program.hpp"
#include "qboot/sdpb_input.hpp"
#include "qboot/xml_input.hpp"

namespace qboot
{
	using algebra::FunctionSymmetry, algebra::Matrix, algebra::Polynomial, algebra::Vector;
	using mp::real;
	using std::move, std::unique_ptr, std::vector;
	using FixedBlock = ConformalBlock<PrimaryOperator>;
	using GeneralBlock = GeneralConformalBlock;

	namespace
	{
		inline int sdpb_digits() { return 3 + int(double(mp::global_prec) * 0.302); }  // src/sdpb_input.cpp:32-35
		inline uint32_t rc_index(uint32_t r, uint32_t c)
		{
			const uint32_t hi = std::max(r, c), lo = std::min(r, c);
			return hi * (hi + 1) / 2 + lo;
		}
		// pool of distinct records (by bit pattern): index of a record, appended on first sight
		class RecordPool
		{
			size_t W_;
			std::unordered_map<std::string, int32_t> index_;

		public:
			std::vector<uint32_t> words;
			explicit RecordPool(size_t W) : W_(W) {}
			int32_t size() const { return int32_t(words.size() / W_); }
			int32_t add(const real& x)
			{
				std::string key(reinterpret_cast<const char*>(x._words()), 4 * W_);
				auto [it, fresh] = index_.emplace(std::move(key), size());
				if (fresh) words.insert(words.end(), x._words(), x._words() + W_);
				return it->second;
			}
		};
		// pool of distinct table keys (Delta, spin, S, P)
		class KeyPool
		{
			size_t W_;
			std::unordered_map<std::string, int32_t> index_;

		public:
			std::vector<uint32_t> delta, S, P, spin;
			explicit KeyPool(size_t W) : W_(W) {}
			int32_t size() const { return int32_t(spin.size()); }
			int32_t add(const real& d, uint32_t sp, const real& s, const real& p)
			{
				std::string key;
				key.reserve(12 * W_ + 4);
				key.append(reinterpret_cast<const char*>(d._words()), 4 * W_);
				key.append(reinterpret_cast<const char*>(&sp), 4);
				key.append(reinterpret_cast<const char*>(s._words()), 4 * W_);
				key.append(reinterpret_cast<const char*>(p._words()), 4 * W_);
				auto [it, fresh] = index_.emplace(std::move(key), size());
				if (fresh)
				{
					delta.insert(delta.end(), d._words(), d._words() + W_);
					S.insert(S.end(), s._words(), s._words() + W_);
					P.insert(P.end(), p._words(), p._words() + W_);
					spin.push_back(sp);
				}
				return it->second;
			}
		};
		real load_record(const uint32_t* w, size_t W)
		{
			real r;
			std::memcpy(r._words(), w, 4 * W);
			return r;
		}
	}  // namespace

	SDPMode::~SDPMode() = default;
	_find_contradiction::~_find_contradiction() = default;
	_extremal_OPE::~_extremal_OPE() = default;
	// src/bootstrap_equation.cpp:340-355
	PolynomialProgram _find_contradiction::_prepare(uint32_t N, const BootstrapEquation& boot) const
	{
		PolynomialProgram prg(N);
		prg.objective_constant() = real(0);
		prg.objectives(Vector<real>(N, real(0)));
		prg.add_equation(boot.make_disc_mat_v(norm_), real(1));
		return prg;
	}
	PolynomialProgram _extremal_OPE::_prepare(uint32_t N, const BootstrapEquation& boot) const
	{
		PolynomialProgram prg(N);
		prg.objective_constant() = real(0);
		prg.objectives(boot.make_disc_mat_v(norm_));
		prg.add_equation(boot.make_disc_mat_v(target_), real(maximize_ ? 1 : -1));
		return prg;
	}
	Vector<real> read_raw_functional(const fs::path& y_txt)
	{
		std::ifstream is(y_txt);
		uint32_t sz = 0, M = 0;
		is >> sz >> M;
		Vector<real> v(sz);
		for (uint32_t i = 0; i < sz; ++i) is >> v.at(i);
		return v;
	}

	// ---- the planner ------------------------------------------------------------------------------------------------------
	// One Plan = a list of constraint blocks (a discrete sector, or one operator of a continuous sector) with their terms,
	// the distinct table keys behind them, and the distinct v-powers and coefficients.
	struct BootstrapEquation::Plan
	{
		struct Item
		{
			uint32_t sector;
			int32_t op_index;                      // -1: discrete sector
			unique_ptr<ConformalScale> scale;      // continuous only
			Vector<real> deltas;                   // sample Deltas (continuous only)
			std::optional<std::array<Matrix<real>, 2>> bilinear{};  // q_m(x_k) sqrt(s_k [x_k])
			std::shared_ptr<void> samples{};                         // BilinearSamples, prepared while the GPU works
		};
		const BootstrapEquation& be;
		std::shared_ptr<b200::Engine> engine;
		size_t W;
		vector<Item> items;
		vector<qb_pmp_block> blocks;
		vector<qb_pmp_term> terms;
		vector<int32_t> tab;
		KeyPool keys;
		RecordPool dvals, coefs;
		uint64_t epoch = 0;
		explicit Plan(const BootstrapEquation& b) : be(b), engine(b.cont_._engine()), W(size_t(engine->words)), keys(W), dvals(W), coefs(W) {}
		void add_discrete(uint32_t id) { items.push_back(Item{id, -1, {}, {}}); }
		void add_continuous(uint32_t id, uint32_t op_index) { items.push_back(Item{id, int32_t(op_index), {}, {}}); }
		// scale factors and sample Deltas of the continuous items, on `parallel` host threads (each is independent)
		void build_scales(uint32_t parallel, const unique_ptr<_event_base>& event)
		{
			vector<std::function<int()>> tasks;
			for (auto& it : items)
				if (it.op_index >= 0 && !it.scale)
					tasks.emplace_back([this, &it, &event] {
						const auto& op = be.sector(it.sector).ops_[size_t(it.op_index)];
						_scoped_event scope(op.str() + " in " + be.sector(it.sector).name(), event);
						// poles and sample points now; the bilinear bases follow underneath the GPU phase (convert)
						it.scale = std::make_unique<ConformalScale>(ConformalScale::_DeferBases{}, op, be.cont_, be.sector_includes_odd(it.sector));
						auto xs = it.scale->sample_points();
						it.deltas = Vector<real>(xs.size());
						for (uint32_t k = 0; k < xs.size(); ++k) it.deltas[k] = it.scale->get_delta(xs[k]);
						return 0;
					});
			_parallel_evaluate(tasks, parallel);
		}
		// the term lists, in the order make_cont_mat / make_disc_mat visit them: equations outer, terms inner
		void build_terms()
		{
			for (auto& it : items)
			{
				const Sector& sec = be.sector(it.sector);
				qb_pmp_block blk{};
				blk.n_samples = it.op_index >= 0 ? int32_t(it.deltas.size()) : 1;
				blk.sz = int32_t(sec.size());
				blk.term0 = int32_t(terms.size());
				for (uint32_t e = 0; e < be.eqs_.size(); ++e)
					for (const auto& term : be.eqs_[e][it.sector])
					{
						qb_pmp_term t{};
						t.eq = int32_t(e);
						const uint32_t r = term.row(), c = term.column();
						t.rc = int32_t(rc_index(r, c));
						t.tab0 = int32_t(tab.size());
						// term.coeff() / (r == c ? 1 : 2), rounded as the reference rounds it (bootstrap_equation.cpp:402,450)
						const real cf = term.coeff() / (r == c ? 1 : 2);
						t.coef = cf == 1 ? -1 : coefs.add(cf);
						if (it.op_index < 0)
						{
							const auto& b = std::get<FixedBlock>(term.block());
							t.d = dvals.add(b.delta_half());
							tab.push_back(keys.add(b.get_op().delta(), b.spin(), b.S(), b.P()));
						}
						else
						{
							const auto& b = std::get<GeneralBlock>(term.block());
							const uint32_t spin = sec.ops_[size_t(it.op_index)].spin();
							t.d = dvals.add(b.delta_half());
							for (uint32_t k = 0; k < it.deltas.size(); ++k) tab.push_back(keys.add(it.deltas[k], spin, b.S(), b.P()));
						}
						terms.push_back(t);
					}
				blk.n_terms = int32_t(terms.size()) - blk.term0;
				blocks.push_back(blk);
			}
		}
		// tables once each, then all blocks: leaves the values in the engine's arena (block i of the arena = items[i])
		void run(const unique_ptr<_event_base>& event = {})
		{
			std::lock_guard<std::recursive_mutex> g(engine->mu);
			vector<int32_t> eq_dim, eq_sym;
			for (const auto& eq : be.eqs_)
			{
				eq_dim.push_back(int32_t(eq.dimension()));
				eq_sym.push_back(int32_t(eq.symmetry()));
			}
			{
				_scoped_event scope("  gblock plan", event);
				engine->check(qb_gblock_plan(engine->h, keys.size(), keys.spin.data(), keys.delta.data(), keys.S.data(), keys.P.data()), "qb_gblock_plan");
			}
			{
				_scoped_event scope("  gblock run", event);
				engine->check(qb_gblock_run(engine->h), "qb_gblock_run");
			}
			_scoped_event scope("  pmp assemble", event);
			engine->check(qb_pmp_assemble(engine->h, int(eq_dim.size()), eq_dim.data(), eq_sym.data(), int(blocks.size()), blocks.data(),
			                              int(terms.size()), terms.data(), int(tab.size()), tab.data(), dvals.size(), dvals.words.data(),
			                              coefs.size(), coefs.words.data()),
			              "qb_pmp_assemble");
			epoch = ++engine->epoch;
		}
		// block i -> mat[n][k](r, c), both triangles filled (the sums for (r, c) and (c, r) are the same numbers)
		Vector<Vector<Matrix<real>>> fetch(size_t i) const
		{
			const auto& blk = blocks[i];
			const uint32_t N = be.N_, ns = uint32_t(blk.n_samples), sz = uint32_t(blk.sz), nrc = sz * (sz + 1) / 2;
			vector<uint32_t> buf(W * size_t(nrc) * ns * N);
			{
				std::lock_guard<std::recursive_mutex> g(engine->mu);
				if (engine->epoch != epoch) throw std::logic_error("qboot_b200: constraint blocks were replaced by a later plan");
				engine->check(qb_pmp_block_fetch(engine->h, int(i), buf.data()), "qb_pmp_block_fetch");
			}
			Vector<Vector<Matrix<real>>> mat(N);
			for (uint32_t n = 0; n < N; ++n)
			{
				mat[n] = Vector<Matrix<real>>(ns);
				for (uint32_t k = 0; k < ns; ++k)
				{
					mat[n][k] = Matrix<real>(sz, sz);
					for (uint32_t r = 0; r < sz; ++r)
						for (uint32_t c = 0; c < sz; ++c)
							mat[n][k].at(r, c) = load_record(buf.data() + W * ((size_t(rc_index(r, c)) * ns + k) * N + n), W);
				}
			}
			return mat;
		}
	};

	// ---- Sector ---------------------------------------------------------------------------------------------------------------
	Sector::Sector(std::string_view name, uint32_t size, SectorType type) : name_(name), sz_(size), type_(type) { assert(sz_ > 0); }
	Sector::Sector(std::string_view name, uint32_t size, Vector<real>&& ope) : name_(name), sz_(size), type_(SectorType::Discrete), ope_{move(ope)}
	{
		assert(sz_ > 0 && ope_->size() == sz_);
	}
	Sector::Sector(const Sector& s) : name_(s.name_), sz_(s.sz_), type_(s.type_), requested_(s.requested_), ops_(s.ops_)
	{
		if (s.ope_) ope_ = s.ope_->clone();
	}
	void Sector::_reset() &&
	{
		name_ = std::string{};
		sz_ = 0;
		type_ = SectorType::Continuous;
		requested_.clear();
		ops_.clear();
		ope_.reset();
	}
	void Sector::request(Request&& r) &
	{
		assert(type_ == SectorType::Continuous);
		requested_.push_back(move(r));
	}
	// src/bootstrap_equation.hpp:64-76 in the reference: the requests become operators, in the order they were made
	void Sector::set_operators(const mp::rational& epsilon, const std::function<uint32_t(uint32_t)>& num_poles) &
	{
		assert(type_ == SectorType::Continuous);
		ops_.clear();
		for (const Request& r : requested_)
		{
			const uint32_t poles = num_poles(r.spin);
			if (!r.from)
				ops_.emplace_back(r.spin, poles, epsilon);
			else if (!r.to)
				ops_.emplace_back(r.spin, poles, epsilon, *r.from);
			else
				ops_.emplace_back(r.spin, poles, epsilon, *r.from, *r.to);
		}
	}

	void BootstrapEquation::finish() &
	{
		for (const auto& eq : eqs_) N_ += eq.dimension();
	}
	bool BootstrapEquation::sector_includes_odd(uint32_t id) const
	{
		for (const auto& eq : eqs_)
			for (const auto& term : eq[id])
				if (std::get<GeneralBlock>(term.block()).include_odd()) return true;
		return false;
	}
	// src/bootstrap_equation.cpp:413-422
	unique_ptr<ConformalScale> BootstrapEquation::common_scale(uint32_t id, const GeneralPrimaryOperator& op) const
	{
		return std::make_unique<ConformalScale>(op, cont_, sector_includes_odd(id));
	}
	Vector<Matrix<real>> BootstrapEquation::make_disc_mat(uint32_t id) const
	{
		assert(sector(id).type() == SectorType::Discrete);
		{
			auto it = disc_cache_.find(id);
			if (it != disc_cache_.end()) return it->second.clone();
		}
		Plan plan(*this);
		plan.add_discrete(id);
		plan.build_terms();
		plan.run();
		auto full = plan.fetch(0);
		Vector<Matrix<real>> mat(N_);
		for (uint32_t n = 0; n < N_; ++n) mat[n] = move(full[n][0]);
		return mat;
	}
	Vector<Vector<Matrix<real>>> BootstrapEquation::make_cont_mat(uint32_t id, const GeneralPrimaryOperator& op,
	                                                               const unique_ptr<ConformalScale>& ag) const
	{
		assert(sector(id).type() == SectorType::Continuous);
		Plan plan(*this);
		// the operator need not be one of the sector's own: plan it as a one-off item
		Plan::Item item{id, 0, {}, {}};
		auto xs = ag->sample_points();
		item.deltas = Vector<real>(xs.size());
		for (uint32_t k = 0; k < xs.size(); ++k) item.deltas[k] = ag->get_delta(xs[k]);
		// build_terms reads the spin from sector.ops_[op_index]; find (or fake) it
		const auto& ops = sector(id).ops_;
		int32_t found = -1;
		for (size_t i = 0; i < ops.size(); ++i)
			if (ops[i].spin() == op.spin()) found = int32_t(i);
		if (found < 0) throw std::invalid_argument("qboot::BootstrapEquation::make_cont_mat: the sector has no operator of this spin");
		item.op_index = found;
		plan.items.push_back(move(item));
		plan.build_terms();
		plan.run();
		return plan.fetch(0);
	}
	Matrix<real> BootstrapEquation::apply_functional(const Vector<real>& func, std::string_view disc_sector) const
	{
		const auto id = get_id(disc_sector);
		auto mats = make_disc_mat(id);
		const auto sz = sector(id).size();
		Matrix<real> mat{sz, sz};
		for (uint32_t n = 0; n < N_; ++n) mat += mul_scalar(func[n], move(mats[n]));
		return mat;
	}
	Vector<real> BootstrapEquation::recover(const unique_ptr<SDPMode>& mode, const Vector<real>& raw_func) const
	{
		assert(N_ > 0);
		return mode->_prepare(N_, *this).recover(raw_func);
	}
	std::map<std::string, std::vector<PrimaryOperator>, std::less<>> BootstrapEquation::get_spectrum(const Vector<real>&, uint32_t,
	                                                                                               const unique_ptr<_event_base>&) const
	{
		throw std::logic_error("qboot_b200: get_spectrum (extremal functional method) is outside the scope of this engine");
	}
	// host half of the bilinear-bases batch for ONE scale factor, built on the host pool: the operands as records, ready to be
	// concatenated into the device call
	struct BilinearSamples
	{
		uint32_t deg = 0;
		vector<uint32_t> coef, xr, sr;  // coefficients of q_0 .. q_{deg/2}; x_k; sqrt(s_k), sqrt(s_k x_k) interleaved
		vector<int32_t> ofs;            // coefficient range of q_m: [ofs[m], ofs[m + 1])
	};
	BilinearSamples prepare_bilinear_samples(const ScaleFactor& chi, size_t W);
	static vector<std::array<Matrix<real>, 2>> bilinear_at_samples_batch(b200::Engine& eng, const vector<const BilinearSamples*>& samples,
	                                                                    uint32_t parallel);
	// src/bootstrap_equation.cpp:555-593, batched
	PolynomialProgram BootstrapEquation::convert(const unique_ptr<SDPMode>& mode, uint32_t parallel, const unique_ptr<_event_base>& event) const
	{
		assert(N_ > 0);
		const auto filter = mode->_get_filter();
		Plan plan(*this);
		// every discrete sector (the mode's _prepare may ask for any of them), and the operators of the continuous sectors
		// that pass the filter -- in the order the reference emits its inequalities: sector names sorted, operators as added
		struct Out
		{
			uint32_t sector;
			size_t item;
		};
		vector<Out> outs;
		for (const auto& [name, id] : sector_id_)
		{
			if (sector(id).type() == SectorType::Discrete)
			{
				if (filter(name)) outs.push_back(Out{id, plan.items.size()});
				plan.add_discrete(id);
			}
			else if (filter(name))
				for (uint32_t i = 0; i < sector(id).ops_.size(); ++i)
				{
					outs.push_back(Out{id, plan.items.size()});
					plan.add_continuous(id, i);
				}
		}
		{
			_scoped_event scope("scale factors", event);
			plan.build_scales(std::max(parallel, cont_.parallel()), event);
		}
		plan.build_terms();
		{
			// The GPU computes the tables and assembles the constraint blocks while the host threads build the bilinear bases of
			// the scale factors (incomplete gamma functions, Hankel-Cholesky: independent of the tables).
			_scoped_event scope("tables + assembly (GPU) | bilinear bases (host)", event);
			std::exception_ptr gpu_error;
			std::thread gpu([&] {
				try
				{
					_scoped_event inner("tables + assembly (GPU)", event);
					plan.run(event);
				}
				catch (...)
				{
					gpu_error = std::current_exception();
				}
			});
			vector<std::function<int()>> tasks;
			const size_t W = plan.W;
			for (auto& it : plan.items)
				if (it.op_index >= 0 && it.scale)
					tasks.emplace_back([&it, W] {
						it.scale->_finish_bases();
						it.samples = std::make_shared<BilinearSamples>(prepare_bilinear_samples(*it.scale, W));
						return 0;
					});
			// largest first (the cost grows with the number of poles, i.e. with the degree): the pool hands tasks out in order, and
			// a long task started last would be the tail of the phase
			{
				vector<std::pair<uint32_t, size_t>> order;
				size_t k = 0;
				for (auto& it : plan.items)
					if (it.op_index >= 0 && it.scale) order.emplace_back(it.scale->max_degree(), k++);
				std::stable_sort(order.begin(), order.end(), [](const auto& a, const auto& b) { return a.first > b.first; });
				vector<std::function<int()>> sorted;
				sorted.reserve(tasks.size());
				for (const auto& o : order) sorted.push_back(move(tasks[o.second]));
				tasks.swap(sorted);
			}
			try
			{
				// one hardware thread is left to the thread that drives the GPU
				_parallel_evaluate(tasks, std::max(1u, std::max(parallel, cont_.parallel()) - 1));
			}
			catch (...)
			{
				gpu.join();
				throw;
			}
			gpu.join();
			if (gpu_error) std::rethrow_exception(gpu_error);
		}
		{
			// the bases at the sample points: one more device batch (create_input's other half, src/polynomial_program.cpp:224-229)
			_scoped_event scope("bilinear bases at the sample points (GPU)", event);
			vector<const BilinearSamples*> smp;
			for (auto& it : plan.items)
				if (it.op_index >= 0 && it.scale) smp.push_back(static_cast<const BilinearSamples*>(it.samples.get()));
			auto bl = bilinear_at_samples_batch(*plan.engine, smp, std::max(parallel, cont_.parallel()));
			size_t i = 0;
			for (auto& it : plan.items)
				if (it.op_index >= 0 && it.scale) it.bilinear = move(bl[i++]);
		}
		// discrete sectors come back to the host: the mode's normalisation / objective vectors are built from them
		disc_cache_.clear();
		for (size_t i = 0; i < plan.items.size(); ++i)
			if (plan.items[i].op_index < 0)
			{
				auto full = plan.fetch(i);
				Vector<Matrix<real>> mat(N_);
				for (uint32_t n = 0; n < N_; ++n) mat[n] = move(full[n][0]);
				disc_cache_.emplace(plan.items[i].sector, move(mat));
			}
		PolynomialProgram prg = mode->_prepare(N_, *this);
		prg._use_engine(plan.engine);
		for (const auto& o : outs)
		{
			const Sector& sec = sector(o.sector);
			auto& item = plan.items[o.item];
			const uint32_t sz = sec.size();
			if (item.op_index >= 0)
			{
				PolynomialInequality ineq(N_, sz, move(item.scale), b200::DeviceBlock{plan.engine, int32_t(o.item), plan.epoch});
				ineq.bilinear_ = move(item.bilinear);
				prg.add_inequality(move(ineq));
			}
			else if (sec.is_matrix())
				prg.add_inequality(PolynomialInequality(N_, sz, std::make_unique<TrivialScale>(), b200::DeviceBlock{plan.engine, int32_t(o.item), plan.epoch}));
			else
				prg.add_inequality(PolynomialInequality(N_, make_disc_mat_v(o.sector), real(0)));
		}
		disc_cache_.clear();
		return prg;
	}

	// ---- PolynomialInequality -----------------------------------------------------------------------------------------------
	// the four host-data constructors of src/polynomial_program.cpp:26-101
	PolynomialInequality::PolynomialInequality(uint32_t N, unique_ptr<ScaleFactor>&& scale, Vector<Vector<real>>&& mat, Vector<real>&& target)
	    : N_(N), sz_(1), chi_(move(scale)), mat_(N)
	{
		const uint32_t deg = chi_->max_degree();
		for (uint32_t i = 0; i < N; ++i)
		{
			assert(mat[i].size() == deg + 1);
			mat_[i] = Vector<Matrix<real>>(deg + 1);
			for (uint32_t k = 0; k <= deg; ++k)
			{
				mat_[i][k] = Matrix<real>(1, 1);
				mat_[i][k].at(0, 0) = mat[i][k];
			}
		}
		target_ = Vector<Matrix<real>>(deg + 1);
		for (uint32_t k = 0; k <= deg; ++k)
		{
			target_[k] = Matrix<real>(1, 1);
			target_[k].at(0, 0) = target[k];
		}
	}
	PolynomialInequality::PolynomialInequality(uint32_t N, Vector<real>&& mat, real&& target)
	    : N_(N), sz_(1), chi_(std::make_unique<TrivialScale>()), mat_(N), target_(1)
	{
		assert(mat.size() == N);
		for (uint32_t i = 0; i < N; ++i)
		{
			mat_[i] = Vector<Matrix<real>>(1);
			mat_[i][0] = Matrix<real>(1, 1);
			mat_[i][0].at(0, 0) = mat[i];
		}
		target_[0] = Matrix<real>(1, 1);
		target_[0].at(0, 0) = target;
	}
	PolynomialInequality::PolynomialInequality(uint32_t N, uint32_t sz, Vector<Matrix<real>>&& mat, Matrix<real>&& target)
	    : N_(N), sz_(sz), chi_(std::make_unique<TrivialScale>()), mat_(N), target_(1)
	{
		assert(mat.size() == N);
		for (uint32_t i = 0; i < N; ++i)
		{
			mat_[i] = Vector<Matrix<real>>(1);
			mat_[i][0] = move(mat[i]);
		}
		target_[0] = move(target);
	}
	PolynomialInequality::PolynomialInequality(uint32_t N, uint32_t sz, unique_ptr<ScaleFactor>&& scale, Vector<Vector<Matrix<real>>>&& mat,
	                                           Vector<Matrix<real>>&& target)
	    : N_(N), sz_(sz), chi_(move(scale)), mat_(move(mat)), target_(move(target))
	{
	}
	PolynomialInequality::PolynomialInequality(uint32_t N, uint32_t sz, unique_ptr<ScaleFactor>&& scale, b200::DeviceBlock dev)
	    : N_(N), sz_(sz), chi_(move(scale)), mat_(), dev_(move(dev))
	{
		const uint32_t deg = chi_->max_degree();
		target_ = Vector<Matrix<real>>(deg + 1);
		for (uint32_t k = 0; k <= deg; ++k) target_[k] = Matrix<real>(sz, sz);
	}
	void PolynomialInequality::_reset() &&
	{
		N_ = sz_ = 0;
		chi_.reset();
		move(mat_)._reset();
		move(target_)._reset();
		bilinear_.reset();
		dev_ = {};
	}
	void PolynomialInequality::fetch() const
	{
		if (mat_.size() == N_ || dev_.block < 0) return;
		const uint32_t ns = max_degree() + 1, nrc = sz_ * (sz_ + 1) / 2;
		auto& e = *dev_.engine;
		const size_t W = size_t(e.words);
		vector<uint32_t> buf(W * size_t(nrc) * ns * N_);
		{
			std::lock_guard<std::recursive_mutex> g(e.mu);
			if (e.epoch != dev_.epoch) throw std::logic_error("qboot_b200: this inequality's device block was replaced by a later convert on the same Context");
			e.check(qb_pmp_block_fetch(e.h, dev_.block, buf.data()), "qb_pmp_block_fetch");
		}
		mat_ = Vector<Vector<Matrix<real>>>(N_);
		for (uint32_t n = 0; n < N_; ++n)
		{
			mat_[n] = Vector<Matrix<real>>(ns);
			for (uint32_t k = 0; k < ns; ++k)
			{
				mat_[n][k] = Matrix<real>(sz_, sz_);
				for (uint32_t r = 0; r < sz_; ++r)
					for (uint32_t c = 0; c < sz_; ++c) mat_[n][k].at(r, c) = load_record(buf.data() + W * ((size_t(rc_index(r, c)) * ns + k) * N_ + n), W);
			}
		}
	}
	// values / chi at the sample points, interpolated (src/polynomial_program.cpp:16-24)
	Matrix<Polynomial> PolynomialInequality::as_polynomial(Vector<Matrix<real>>&& vals)
	{
		const uint32_t deg = vals.size() - 1;
		auto xs = chi_->sample_points();
		const auto inter = algebra::interpolation_matrix(xs);
		Vector<Matrix<real>> ev(deg + 1);
		for (uint32_t k = 0; k <= deg; ++k) ev[k] = vals[k] / chi_->eval(xs[k]);
		Matrix<Polynomial> ans(sz_, sz_);
		Vector<real> v(deg + 1);
		for (uint32_t r = 0; r < sz_; ++r)
			for (uint32_t c = 0; c < sz_; ++c)
			{
				for (uint32_t k = 0; k <= deg; ++k) v[k] = ev[k].at(r, c);
				ans.at(r, c) = algebra::polynomial_interpolate(v, inter);
			}
		return ans;
	}

	// ---- PolynomialProgram ----------------------------------------------------------------------------------------------------
	// src/polynomial_program.cpp:103-116
	Vector<real> PolynomialProgram::recover(const Vector<real>& z)
	{
		const uint32_t eq_sz = uint32_t(eq_rows_.size()), M = N_ - eq_sz;
		assert(z.size() == M);
		Vector<real> y(N_);
		for (uint32_t n = 0; n < M; ++n) y[free_cols_[n]] = z[n];
		for (uint32_t e = 0; e < eq_sz; ++e)
		{
			const auto n = pivot_cols_[e];
			y[n] = eq_rhs_[e];
			for (uint32_t m = 0; m < M; ++m) y[n] -= eq_rows_[e][free_cols_[m]] * z[m];
		}
		return y;
	}
	// Gaussian elimination step with the largest entry as pivot (src/polynomial_program.cpp:118-151)
	PolynomialProgram::PolynomialProgram(uint32_t num_of_vars) : N_(num_of_vars), objective_(num_of_vars), free_cols_(num_of_vars)
	{
		std::iota(free_cols_.begin(), free_cols_.end(), 0u);
	}
	void PolynomialProgram::_reset() &&
	{
		N_ = 0;
		move(objective_offset_)._reset();
		move(objective_)._reset();
		const auto release = [](auto& v) { std::decay_t<decltype(v)>().swap(v); };  // clear() would keep the capacity
		release(eq_rows_);
		release(eq_rhs_);
		release(pivot_cols_);
		release(free_cols_);
		release(ineqs_);
		engine_.reset();
	}
	uint32_t PolynomialProgram::_total_memory() const
	{
		uint32_t held = 1 + objective_.size();
		for (const auto& q : ineqs_)
			if (q.has_value()) held += q->_total_memory();
		for (const auto& row : eq_rows_) held += row.size() + 1;
		return held;
	}
	void PolynomialProgram::objectives(Vector<real>&& obj) &
	{
		assert(obj.size() == N_);
		objective_ = move(obj);
	}
	void PolynomialProgram::add_inequality(std::optional<PolynomialInequality>&& ineq) &
	{
		assert(ineq.has_value() && ineq->num_of_variables() == N_);
		ineqs_.push_back(move(ineq));
	}
	void PolynomialProgram::add_equation(Vector<real>&& vec, real&& target) &
	{
		assert(vec.size() == N_);
		// reduce the new equation by the earlier ones
		for (uint32_t e = 0; e < eq_rows_.size(); ++e)
		{
			const real t = vec[pivot_cols_[e]];
			target -= t * eq_rhs_[e];
			vec -= mul_scalar(t, eq_rows_[e]);
		}
		uint32_t pivot = 0;
		for (uint32_t n = 1; n < N_; ++n)
			if (mp::cmpabs(vec[pivot], vec[n]) < 0) pivot = n;
		const real top = vec[pivot];
		target /= top;
		vec /= top;
		vec[pivot] = 1;
		// and the earlier ones by the new equation
		for (uint32_t e = 0; e < eq_rows_.size(); ++e)
		{
			const real t = eq_rows_[e][pivot];
			eq_rhs_[e] -= t * target;
			eq_rows_[e] -= mul_scalar(t, vec);
		}
		eq_rows_.push_back(move(vec));
		eq_rhs_.push_back(move(target));
		pivot_cols_.push_back(pivot);
		free_cols_.erase(std::find(free_cols_.begin(), free_cols_.end(), pivot));
	}
	// objective in terms of the free variables (src/polynomial_program.cpp:162-170)
	std::pair<real, Vector<real>> PolynomialProgram::eliminated_objective()
	{
		const uint32_t eq_sz = uint32_t(eq_rows_.size()), M = N_ - eq_sz;
		Vector<real> obj_new(M);
		for (uint32_t m = 0; m < M; ++m) obj_new[m] = objective_[free_cols_[m]];
		real constant = objective_offset_;
		for (uint32_t e = 0; e < eq_sz; ++e)
		{
			const real t = objective_[pivot_cols_[e]];
			for (uint32_t m = 0; m < M; ++m) obj_new[m] -= eq_rows_[e][free_cols_[m]] * t;
			constant += eq_rhs_[e] * t;
		}
		return {move(constant), move(obj_new)};
	}
	// bilinear bases at the sample points (src/polynomial_program.cpp:224-229): q0[m][k] = q_m(x_k) sqrt(s_k),
	// q1[m][k] = q_m(x_k) sqrt(s_k x_k) -- for any number of scale factors in ONE device batch (qb_poly_eval_batch: fma Horner
	// and the final product on the GPU; the square roots, 2 (D + 1) per constraint, stay on the host)
	// host half of that batch, one scale factor at a time (thread-safe: run on the host pool): the sample points and the
	// two square roots per point
	BilinearSamples prepare_bilinear_samples(const ScaleFactor& chi, size_t W)
	{
		BilinearSamples in;
		const auto put = [W](vector<uint32_t>& v, const real& r) { v.insert(v.end(), r._words(), r._words() + W); };
		const auto xs = chi.sample_points();
		const auto scs = chi.sample_scalings();
		const auto q = chi.bilinear_bases();
		in.deg = chi.max_degree();
		in.xr.reserve(W * xs.size());
		in.sr.reserve(2 * W * xs.size());
		for (uint32_t k = 0; k < xs.size(); ++k)
		{
			put(in.xr, xs[k]);
			put(in.sr, mp::sqrt(scs[k]));
			put(in.sr, mp::sqrt(scs[k] * xs[k]));
		}
		in.ofs.push_back(0);
		for (uint32_t m = 0; m <= in.deg / 2; ++m)
		{
			for (int32_t p = 0; p <= q[m].degree(); ++p) put(in.coef, q[m].at(uint32_t(p)));
			in.ofs.push_back(int32_t(in.coef.size() / W));
		}
		return in;
	}
	// the device batch over any number of prepared scale factors; concatenation and the call are serial (block copies), the
	// results are turned into matrices on `parallel` host threads
	static vector<std::array<Matrix<real>, 2>> bilinear_at_samples_batch(b200::Engine& eng, const vector<const BilinearSamples*>& samples,
	                                                                    uint32_t parallel)
	{
		const size_t W = size_t(eng.words);
		vector<int32_t> ofs{0};
		vector<uint32_t> coef, xr, sr;
		vector<qb_eval_item> items;
		vector<size_t> first(samples.size());  // first result of each scale factor
		{
			size_t nc = 0, nx = 0, ni = 0;
			for (const BilinearSamples* in : samples)
			{
				nc += in->coef.size();
				nx += in->xr.size();
				const size_t d0 = in->deg / 2, d1 = in->deg == 0 ? 0 : (in->deg - 1) / 2;
				ni += (d0 + 1 + d1 + 1) * (size_t(in->deg) + 1);
			}
			coef.reserve(nc);
			xr.reserve(nx);
			sr.reserve(2 * nx);
			items.reserve(ni);
		}
		for (size_t ci = 0; ci < samples.size(); ++ci)
		{
			const BilinearSamples& in = *samples[ci];
			const uint32_t deg = in.deg, d0 = deg / 2, d1 = deg == 0 ? 0 : (deg - 1) / 2;
			const int32_t pb = int32_t(ofs.size()) - 1, cb = int32_t(coef.size() / W), xb = int32_t(xr.size() / W), sb = int32_t(sr.size() / W);
			coef.insert(coef.end(), in.coef.begin(), in.coef.end());
			xr.insert(xr.end(), in.xr.begin(), in.xr.end());
			sr.insert(sr.end(), in.sr.begin(), in.sr.end());
			for (size_t m = 1; m < in.ofs.size(); ++m) ofs.push_back(cb + in.ofs[m]);
			first[ci] = items.size();
			for (uint32_t m = 0; m <= d0; ++m)
				for (uint32_t k = 0; k <= deg; ++k) items.push_back(qb_eval_item{pb + int32_t(m), xb + int32_t(k), sb + 2 * int32_t(k)});
			for (uint32_t m = 0; m <= d1; ++m)
				for (uint32_t k = 0; k <= deg; ++k) items.push_back(qb_eval_item{pb + int32_t(m), xb + int32_t(k), sb + 2 * int32_t(k) + 1});
		}
		vector<uint32_t> out(W * items.size());
		if (!items.empty())
		{
			std::lock_guard<std::recursive_mutex> g(eng.mu);
			eng.check(qb_poly_eval_batch(eng.h, int(ofs.size()) - 1, ofs.data(), coef.data(), int(xr.size() / W), xr.data(), int(sr.size() / W),
			                             sr.data(), int(items.size()), items.data(), out.data()),
			          "qb_poly_eval_batch");
		}
		vector<std::array<Matrix<real>, 2>> res(samples.size());
		vector<std::function<int()>> tasks;
		for (size_t ci = 0; ci < samples.size(); ++ci)
			tasks.emplace_back([&, ci] {
				const uint32_t deg = samples[ci]->deg, d0 = deg / 2, d1 = deg == 0 ? 0 : (deg - 1) / 2;
				size_t pos = first[ci];
				Matrix<real> q0(d0 + 1, deg + 1), q1(d1 + 1, deg + 1);
				for (uint32_t m = 0; m <= d0; ++m)
					for (uint32_t k = 0; k <= deg; ++k, ++pos) std::memcpy(q0.at(m, k)._words(), out.data() + W * pos, 4 * W);
				for (uint32_t m = 0; m <= d1; ++m)
					for (uint32_t k = 0; k <= deg; ++k, ++pos) std::memcpy(q1.at(m, k)._words(), out.data() + W * pos, 4 * W);
				res[ci] = {move(q0), move(q1)};
				return 0;
			});
		_parallel_evaluate(tasks, std::max(1u, parallel));
		return res;
	}
	// src/polynomial_program.cpp:152-232 -- the per-inequality loops run as ONE pack kernel over all blocks
	SDPBInput PolynomialProgram::create_input(uint32_t parallel, const unique_ptr<_event_base>& event) &&
	{
		const uint32_t eq_sz = uint32_t(eq_rows_.size()), M = N_ - eq_sz, J = uint32_t(ineqs_.size());
		auto [constant, obj_new] = eliminated_objective();
		SDPBInput sdpb(move(constant), move(obj_new), J);
		// the engine that holds the device-resident blocks (or a fresh one for a program built from host data only)
		std::shared_ptr<b200::Engine> eng = engine_;
		for (const auto& q : ineqs_)
			if (!eng && q->_on_device()) eng = q->dev_.engine;
		if (!eng) eng = std::make_shared<b200::Engine>(mp::_prec(), 1u, 1u, mp::rational(1, 2));
		const size_t W = size_t(eng->words);
		auto packed = std::make_shared<b200::PackedProgram>();
		packed->engine = eng;
		packed->n_free = int(M);
		packed->n_blocks = int(J);
		{
			_scoped_event scope("pack (GPU)", event);
			std::lock_guard<std::recursive_mutex> g(eng->mu);
			int64_t info[4] = {0, 0, 0, 0};
			eng->check(qb_pmp_info(eng->h, info), "qb_pmp_info");
			if (info[1] != int64_t(N_))
			{
				// nothing of this program is on the device yet: open an empty arena with N variables
				for (const auto& q : ineqs_)
					if (q->_on_device()) throw std::logic_error("qboot_b200: device blocks do not belong to this engine's current program");
				const int32_t dim = int32_t(N_), sym = 1;
				eng->check(qb_pmp_assemble(eng->h, 1, &dim, &sym, 0, nullptr, 0, nullptr, 0, nullptr, 0, nullptr, 0, nullptr), "qb_pmp_assemble");
				++eng->epoch;
			}
			vector<int32_t> sel(J);
			for (uint32_t j = 0; j < J; ++j)
			{
				auto& q = ineqs_[j].value();
				if (q._on_device())
				{
					if (q.dev_.engine != eng || q.dev_.epoch != eng->epoch)
						throw std::logic_error("qboot_b200: an inequality's device block was replaced by a later convert on the same Context");
					sel[j] = q.dev_.block;
					continue;
				}
				// host data: upload in the arena's layout M[p][n], p = (rc, k), and the target values
				const uint32_t ns = q.max_degree() + 1, sz = q.size(), nrc = sz * (sz + 1) / 2;
				vector<uint32_t> mat(W * size_t(nrc) * ns * N_), tgt(W * size_t(nrc) * ns);
				for (uint32_t r = 0; r < sz; ++r)
					for (uint32_t c = 0; c <= r; ++c)
						for (uint32_t k = 0; k < ns; ++k)
						{
							const size_t p = size_t(rc_index(r, c)) * ns + k;
							for (uint32_t n = 0; n < N_; ++n) std::memcpy(mat.data() + W * (p * N_ + n), q.mat_[n][k].at(r, c)._words(), 4 * W);
							std::memcpy(tgt.data() + W * p, q.target_[k].at(r, c)._words(), 4 * W);
						}
				sel[j] = eng->check(qb_pmp_block_add(eng->h, int(ns), int(sz), mat.data(), tgt.data()), "qb_pmp_block_add", true);
			}
			vector<int32_t> lead(pivot_cols_.begin(), pivot_cols_.end()), free_idx(free_cols_.begin(), free_cols_.end());
			vector<uint32_t> eqn(W * size_t(eq_sz) * N_), eqt(W * size_t(eq_sz));
			for (uint32_t e = 0; e < eq_sz; ++e)
			{
				for (uint32_t n = 0; n < N_; ++n) std::memcpy(eqn.data() + W * (size_t(e) * N_ + n), eq_rows_[e][n]._words(), 4 * W);
				std::memcpy(eqt.data() + W * e, eq_rhs_[e]._words(), 4 * W);
			}
			eng->check(qb_pmp_pack(eng->h, int(J), sel.data(), int(eq_sz), lead.data(), int(M), free_idx.data(), eqn.data(), eqt.data()), "qb_pmp_pack");
			packed->pack_epoch = ++eng->pack_epoch;
		}
		// bilinear bases at the sample points: convert has evaluated them for the constraints it built; the others (inequalities
		// from host data) go through one device batch now
		{
			vector<const ScaleFactor*> chis;
			vector<uint32_t> which;
			for (uint32_t j = 0; j < J; ++j)
				if (!ineqs_[j]->bilinear_)
				{
					chis.push_back(ineqs_[j]->chi_.get());
					which.push_back(j);
				}
			vector<BilinearSamples> prepared(chis.size());
			vector<std::function<int()>> prep;
			for (size_t i = 0; i < chis.size(); ++i)
				prep.emplace_back([&prepared, &chis, i, W] {
					prepared[i] = prepare_bilinear_samples(*chis[i], W);
					return 0;
				});
			_parallel_evaluate(prep, parallel);
			vector<const BilinearSamples*> smp;
			for (const auto& x : prepared) smp.push_back(&x);
			auto bl = bilinear_at_samples_batch(*eng, smp, parallel);
			for (size_t i = 0; i < which.size(); ++i) ineqs_[which[i]]->bilinear_ = move(bl[i]);
		}
		vector<std::function<int()>> tasks;
		for (uint32_t j = 0; j < J; ++j)
			tasks.emplace_back([this, &sdpb, &packed, j, M, &event] {
				_scoped_event scope(std::to_string(j), event);
				auto& ineq = ineqs_[j].value();
				const uint32_t sz = ineq.size(), deg = ineq.max_degree();
				sdpb.register_constraint(j, DualConstraint(sz, deg, packed, int32_t(j), int32_t(M), move(*ineq.bilinear_)));
				ineqs_[j].reset();
				return 0;
			});
		_parallel_evaluate(tasks, parallel);
		eq_rows_.clear();
		eq_rhs_.clear();
		free_cols_.clear();
		ineqs_.clear();
		pivot_cols_.clear();
		return sdpb;
	}
	// src/polynomial_program.cpp:233-281 (host; legacy output)
	XMLInput PolynomialProgram::create_xml(uint32_t parallel, const unique_ptr<_event_base>& event) &&
	{
		const uint32_t eq_sz = uint32_t(eq_rows_.size()), M = N_ - eq_sz, J = uint32_t(ineqs_.size());
		auto [constant, obj_new] = eliminated_objective();
		XMLInput xml(move(constant), move(obj_new), J);
		vector<std::function<int()>> tasks;
		for (uint32_t j = 0; j < J; ++j)
			tasks.emplace_back([this, &xml, j, M, eq_sz, &event] {
				_scoped_event scope(std::to_string(j), event);
				auto& ineq = ineqs_[j].value();
				const uint32_t sz = ineq.size();
				Matrix<Vector<Polynomial>> mat(sz, sz);
				for (uint32_t r = 0; r < sz; ++r)
					for (uint32_t c = 0; c < sz; ++c) mat.at(r, c) = Vector<Polynomial>(M + 1);
				Matrix<Polynomial> target = -ineq.target_polynomial();
				Vector<Matrix<Polynomial>> new_mat(M);
				for (uint32_t m = 0; m < M; ++m) new_mat[m] = ineq.matrix_polynomial(free_cols_[m]);
				for (uint32_t e = 0; e < eq_sz; ++e)
				{
					auto t = ineq.matrix_polynomial(pivot_cols_[e]);
					for (uint32_t m = 0; m < M; ++m) new_mat[m] -= mul_scalar(eq_rows_[e][free_cols_[m]], t);
					target += mul_scalar(eq_rhs_[e], t);
				}
				for (uint32_t r = 0; r < sz; ++r)
					for (uint32_t c = 0; c < sz; ++c)
					{
						mat.at(r, c).at(0) = move(target.at(r, c));
						for (uint32_t m = 0; m < M; ++m) mat.at(r, c).at(m + 1) = move(new_mat[m].at(r, c));
					}
				xml.register_constraint(j, PVM(move(mat), ineq.sample_points(), ineq.sample_scalings(), ineq.bilinear_bases()));
				ineqs_[j].reset();
				return 0;
			});
		_parallel_evaluate(tasks, parallel);
		eq_rows_.clear();
		eq_rhs_.clear();
		free_cols_.clear();
		ineqs_.clear();
		pivot_cols_.clear();
		return xml;
	}

	// ---- SDPB input ---------------------------------------------------------------------------------------------------------
	namespace b200
	{
		// text of every d_B and d_c number of the packed program, produced on the device
		void PackedProgram::ensure_text(int ndigits)
		{
			std::lock_guard<std::mutex> own(mu);
			if (text_digits == ndigits && (text || text_bytes == 0) && !sizes.empty()) return;
			std::lock_guard<std::recursive_mutex> g(engine->mu);
			require_current();
			const size_t W = size_t(engine->words);
			sizes.assign(size_t(2 * n_blocks), 0);
			int declined = engine->check(qb_pmp_text(engine->h, ndigits, sizes.data()), "qb_pmp_text", true);
			if (declined > 0)
			{
				// exponents outside the device formatter's range: print those on the host and hand the text back
				vector<int64_t> idx(static_cast<size_t>(declined));
				vector<uint32_t> recs(W * size_t(declined));
				const int got = engine->check(qb_pmp_text_declined(engine->h, declined, idx.data(), recs.data()), "qb_pmp_text_declined", true);
				const int stride = ndigits + 16;
				vector<char> texts(size_t(got) * size_t(stride));
				vector<int32_t> len(static_cast<size_t>(got));
				for (int i = 0; i < got; ++i)
				{
					std::string s = mp::_format(load_record(recs.data() + W * size_t(i), W), std::ios_base::fmtflags{}, ndigits) + "\n";
					if (int(s.size()) > stride) throw std::runtime_error("qboot_b200: number too long for its text slot");
					std::memcpy(texts.data() + size_t(i) * size_t(stride), s.data(), s.size());
					len[size_t(i)] = int32_t(s.size());
				}
				engine->check(qb_pmp_text_patch(engine->h, got, idx.data(), texts.data(), stride, len.data(), sizes.data()), "qb_pmp_text_patch");
			}
			offsets.assign(sizes.size() + 1, 0);
			for (size_t i = 0; i < sizes.size(); ++i) offsets[i + 1] = offsets[i] + sizes[i];
			text_bytes = offsets.back();
			if (size_t(text_bytes) > text_cap)
			{
				pinned_release(text, text_cap);
				text = nullptr;
				text_cap = 0;
				text = pinned_acquire(size_t(text_bytes), text_cap);
			}
			engine->check(qb_pmp_text_fetch(engine->h, text), "qb_pmp_text_fetch");
			text_digits = ndigits;
		}
	}  // namespace b200

	DualConstraint::DualConstraint(uint32_t dim, uint32_t deg, Matrix<real>&& B, Vector<real>&& c, std::array<Matrix<real>, 2>&& bilinear)
	    : shape_{dim, deg}, host_B_(move(B)), host_c_(move(c)), bilinear_(move(bilinear))
	{
		assert(host_B_.row() == schur_size() && host_c_.size() == schur_size());
	}
	DualConstraint::DualConstraint(uint32_t dim, uint32_t deg, std::shared_ptr<b200::PackedProgram> dev, int32_t slot, int32_t cols,
	                               std::array<Matrix<real>, 2>&& bilinear)
	    : shape_{dim, deg}, bilinear_(move(bilinear)), dev_(move(dev)), slot_(slot), cols_(cols)
	{
	}
	void DualConstraint::_reset() &&
	{
		*this = DualConstraint{};
	}
	uint32_t DualConstraint::_total_memory() const
	{
		uint32_t held = schur_size() * (num_free() + 1);
		for (const auto& q : bilinear_) held += q.row() * q.column();
		return held;
	}
	void DualConstraint::fetch() const
	{
		if (!dev_ || host_c_.size() == schur_size()) return;
		auto& e = *dev_->engine;
		const size_t W = size_t(e.words), schur = schur_size(), M = size_t(cols_);
		vector<uint32_t> B(W * schur * M), c(W * schur);
		{
			std::lock_guard<std::recursive_mutex> g(e.mu);
			dev_->require_current();
			e.check(qb_pmp_packed_fetch(e.h, slot_, B.data(), c.data()), "qb_pmp_packed_fetch");
		}
		host_B_ = Matrix<real>(uint32_t(schur), uint32_t(M));
		host_c_ = Vector<real>(uint32_t(schur));
		for (size_t p = 0; p < schur; ++p)
		{
			for (size_t m = 0; m < M; ++m) host_B_.at(uint32_t(p), uint32_t(m)) = load_record(B.data() + W * (p * M + m), W);
			host_c_[uint32_t(p)] = load_record(c.data() + W * p, W);
		}
	}
	SDPBInput::SDPBInput(real&& constant, Vector<real>&& obj, uint32_t num_constraints)
	    : offset_(move(constant)), weights_(move(obj)), slots_(num_constraints)
	{
	}
	void SDPBInput::_reset() &&
	{
		move(offset_)._reset();
		move(weights_)._reset();
		slots_.clear();
	}
	uint32_t SDPBInput::_total_memory() const
	{
		uint32_t held = 1 + weights_.size();
		for (const auto& c : slots_)
			if (c.has_value()) held += c->_total_memory();
		return held;
	}
	void SDPBInput::register_constraint(uint32_t index, DualConstraint&& c) &
	{
		assert(index < slots_.size() && !slots_[index].has_value());
		assert(weights_.size() == c.num_free());
		slots_[index] = move(c);
	}
	namespace
	{
		void append_records(vector<uint32_t>& buf, const Matrix<real>& m, size_t W)
		{
			for (uint32_t r = 0; r < m.row(); ++r)
				for (uint32_t c = 0; c < m.column(); ++c) buf.insert(buf.end(), m.at(r, c)._words(), m.at(r, c)._words() + W);
		}
		void write_file(const fs::path& p, const std::string& head, const char* body, size_t n)
		{
			std::ofstream f(p, std::ios::binary);
			f.write(head.data(), std::streamsize(head.size()));
			if (n) f.write(body, std::streamsize(n));
			if (!f) throw std::runtime_error("qboot_b200: cannot write " + p.string());
		}
	}  // namespace
	// Files as src/sdpb_input.cpp:67-137 lays them out.  Numbers: the large arrays come as text from the device; the rest
	// (objectives, bilinear bases of all constraints) is formatted there too in one small batch.
	void SDPBInput::write_all(const fs::path& root_, uint32_t parallel, const unique_ptr<_event_base>& event) const
	{
		const auto root = root_ / "";
		fs::create_directory(root);
		const int nd = sdpb_digits();
		const uint32_t J = num_constraints();
		// the packed program shared by the device-resident constraints
		std::shared_ptr<b200::PackedProgram> packed;
		for (uint32_t j = 0; j < J; ++j)
			if (slots_[j]->dev_) packed = slots_[j]->dev_;
		// the device formats the large arrays while host threads write the small files (blocks, objectives, bilinear bases)
		// One overlapped phase, three kinds of workers:
		//   * the text thread: the device formats the large arrays (d_B, d_c of every constraint) -> text_ready;
		//   * the host pool: blocks.0, objectives, the text of bilinear_bases.0 piece by piece (the device formats each piece's
		//     numbers) -> piece_ready[i]; then the per-constraint files, which wait for text_ready;
		//   * the writer thread: bilinear_bases.0 is ONE sequential file (a third of the single-correlator problem's text):
		//     piece i is written as soon as it is ready, underneath the formatting of the later pieces.
		std::mutex mu;
		std::condition_variable cv;
		bool text_ready = !packed, failed = false;
		vector<char> piece_ready(J, 0);
		const auto signal = [&](auto&& change) {
			{
				std::lock_guard<std::mutex> g(mu);
				change();
			}
			cv.notify_all();
		};
		const auto wait_for = [&](auto&& ready) {
			std::unique_lock<std::mutex> lk(mu);
			cv.wait(lk, [&] { return failed || ready(); });
			return !failed;
		};
		std::exception_ptr text_error, writer_error;
		std::thread text_thread, writer_thread;
		struct Joiner
		{
			std::thread& t;
			~Joiner()
			{
				if (t.joinable()) t.join();
			}
		};
		if (packed)
			text_thread = std::thread([&] {
				try
				{
					_scoped_event scope("decimal text (GPU)", event);
					packed->ensure_text(nd);
					signal([&] { text_ready = true; });
				}
				catch (...)
				{
					text_error = std::current_exception();
					signal([&] { failed = true; });
				}
			});
		Joiner join_text{text_thread};
		std::shared_ptr<b200::Engine> eng = packed ? packed->engine : std::make_shared<b200::Engine>(mp::_prec(), 1u, 1u, mp::rational(1, 2));
		const size_t W = size_t(eng->words);
		vector<std::string> bl_pieces(J);
		writer_thread = std::thread([&] {
			try
			{
				_scoped_event scope("write_bilinear_bases", event);
				std::ofstream f(root / "bilinear_bases.0", std::ios::binary);
				const std::string head = std::to_string(J) + "\n";
				f.write(head.data(), std::streamsize(head.size()));
				for (uint32_t i = 0; i < J; ++i)
				{
					if (!wait_for([&] { return piece_ready[i] != 0; })) return;
					f.write(bl_pieces[i].data(), std::streamsize(bl_pieces[i].size()));
					std::string().swap(bl_pieces[i]);
				}
				if (!f) throw std::runtime_error("qboot_b200: cannot write " + (root / "bilinear_bases.0").string());
			}
			catch (...)
			{
				writer_error = std::current_exception();
				signal([&] { failed = true; });
			}
		});
		Joiner join_writer{writer_thread};
		vector<std::function<int()>> tasks;
		for (uint32_t i = 0; i < J; ++i)
			tasks.emplace_back([&, i] {
				vector<uint32_t> recs;
				for (const auto& m : slots_[i]->bilinear()) append_records(recs, m, W);
				std::string text;
				vector<size_t> ends;
				b200::format_numbers(*eng, nd, recs.data(), recs.size() / W, text, &ends);
				std::string& out = bl_pieces[i];
				out.reserve(text.size() + 64);
				size_t pos = 0, num = 0;
				for (const auto& m : slots_[i]->bilinear())
				{
					out += std::to_string(m.row()) + " " + std::to_string(m.column()) + "\n";
					num += size_t(m.row()) * m.column();
					const size_t end = num ? ends[num - 1] : 0;
					out.append(text, pos, end - pos);
					pos = end;
				}
				signal([&] { piece_ready[i] = 1; });
				return 0;
			});
		tasks.emplace_back([&] {
			_scoped_event scope("write_blocks", event);
			std::ostringstream out;
			out << 1 << "\n" << J << "\n";
			for (uint32_t i = 0; i < J; ++i) out << i << "\n";
			out << J << "\n";
			for (uint32_t i = 0; i < J; ++i) out << slots_[i]->dim() << "\n";
			out << J << "\n";
			for (uint32_t i = 0; i < J; ++i) out << slots_[i]->degree() << "\n";
			out << J << "\n";
			for (uint32_t i = 0; i < J; ++i) out << slots_[i]->schur_size() << "\n";
			out << 2 * J << "\n";
			for (uint32_t i = 0; i < J; ++i)
				for (const auto& m : slots_[i]->bilinear()) out << m.row() * slots_[i]->dim() << "\n";
			out << 2 * J << "\n";
			for (uint32_t i = 0; i < J; ++i)
				for (const auto& m : slots_[i]->bilinear()) out << m.column() * slots_[i]->dim() << "\n";
			write_file(root / "blocks.0", out.str(), nullptr, 0);
			return 0;
		});
		tasks.emplace_back([&] {
			_scoped_event scope("write_objectives", event);
			vector<uint32_t> recs(offset_._words(), offset_._words() + W);
			for (const auto& x : weights_) recs.insert(recs.end(), x._words(), x._words() + W);
			std::string text;
			b200::format_numbers(*eng, nd, recs.data(), recs.size() / W, text);
			const size_t first = text.find('\n') + 1;
			std::string out = text.substr(0, first) + std::to_string(weights_.size()) + "\n";
			out.append(text, first, std::string::npos);
			write_file(root / "objectives", out, nullptr, 0);
			return 0;
		});
		for (uint32_t i = 0; i < J; ++i)
			tasks.emplace_back([&, i] {
				if (!wait_for([&] { return text_ready; })) return 0;
				_scoped_event scope("write constraint " + std::to_string(i), event);
				const auto& c = slots_[i].value();
				const std::string hb = std::to_string(c.schur_size()) + " " + std::to_string(c.num_free()) + "\n", hc = std::to_string(c.schur_size()) + "\n";
				if (c.dev_)
				{
					const auto& pk = *c.dev_;
					const size_t s = size_t(c.slot_);
					write_file(root / ("free_var_matrix." + std::to_string(i)), hb, pk.text + pk.offsets[2 * s], size_t(pk.sizes[2 * s]));
					write_file(root / ("primal_objective_c." + std::to_string(i)), hc, pk.text + pk.offsets[2 * s + 1], size_t(pk.sizes[2 * s + 1]));
				}
				else
				{
					vector<uint32_t> recs;
					append_records(recs, c.host_B_, W);
					std::string tb, tc;
					b200::format_numbers(*eng, nd, recs.data(), recs.size() / W, tb);
					recs.clear();
					for (const auto& x : c.host_c_) recs.insert(recs.end(), x._words(), x._words() + W);
					b200::format_numbers(*eng, nd, recs.data(), recs.size() / W, tc);
					write_file(root / ("free_var_matrix." + std::to_string(i)), hb, tb.data(), tb.size());
					write_file(root / ("primal_objective_c." + std::to_string(i)), hc, tc.data(), tc.size());
				}
				return 0;
			});
		try
		{
			_scoped_event scope("write: host pool", event);
			_parallel_evaluate(tasks, parallel);
		}
		catch (...)
		{
			signal([&] { failed = true; });
			throw;  // the joiners wait for the two threads
		}
		if (text_thread.joinable()) text_thread.join();
		if (writer_thread.joinable()) writer_thread.join();
		if (text_error) std::rethrow_exception(text_error);
		if (writer_error) std::rethrow_exception(writer_error);
	}
	void SDPBInput::write(const fs::path& root, uint32_t parallel, const unique_ptr<_event_base>& event) const& { write_all(root, parallel, event); }
	void SDPBInput::write(const fs::path& root, uint32_t parallel, const unique_ptr<_event_base>& event) &&
	{
		write_all(root, parallel, event);
		move(weights_)._reset();
		for (auto& c : slots_) c.reset();
	}

	// ---- XML (src/xml_input.cpp:80-138) --------------------------------------------------------------------------------------
	XMLInput::XMLInput(real&& constant, Vector<real>&& obj, uint32_t num_constraints) : objective_row_(obj.size() + 1), slots_(num_constraints)
	{
		objective_row_[0] = move(constant);
		for (uint32_t i = 0; i < obj.size(); ++i) objective_row_[i + 1] = obj[i];
	}
	void XMLInput::register_constraint(uint32_t index, PVM&& c) &
	{
		assert(index < slots_.size() && !slots_[index].has_value());
		slots_[index] = move(c);
	}
	void XMLInput::write(const fs::path& path) const
	{
		std::ofstream f(path);
		f << std::defaultfloat;
		f.precision(sdpb_digits());
		auto open = [&](const char* tag) { f << "<" << tag << ">\n"; };
		auto close = [&](const char* tag) { f << "</" << tag << ">\n"; };
		auto vec = [&](const Vector<real>& v, const char* tag) {
			open(tag);
			for (const auto& x : v)
			{
				open("elt");
				f << x;
				close("elt");
			}
			close(tag);
		};
		auto pol = [&](const Polynomial& p) {
			open("polynomial");
			if (p.iszero())
			{
				open("coeff");
				f << 0;
				close("coeff");
			}
			else
				for (const auto& x : p)
				{
					open("coeff");
					f << x;
					close("coeff");
				}
			close("polynomial");
		};
		open("sdp");
		vec(objective_row_, "objective");
		open("polynomialVectorMatrices");
		for (size_t j = 0; j < slots_.size(); ++j)
		{
			const auto& c = slots_[j].value();
			open("polynomialVectorMatrix");
			open("rows");
			f << c.dim();
			close("rows");
			open("cols");
			f << c.dim();
			close("cols");
			open("elements");
			for (uint32_t r = 0; r < c.dim(); ++r)
				for (uint32_t s = 0; s < c.dim(); ++s)
				{
					open("polynomialVector");
					for (uint32_t n = 0; n <= c.num_of_vars(); ++n) pol(c.matrices().at(r, s).at(n));
					close("polynomialVector");
				}
			close("elements");
			vec(c.sample_points(), "samplePoints");
			vec(c.sample_scalings(), "sampleScalings");
			open("bilinearBasis");
			for (const auto& q : c.bilinear()) pol(q);
			close("bilinearBasis");
			close("polynomialVectorMatrix");
		}
		close("polynomialVectorMatrices");
		close("sdp");
	}
}  // namespace qboot
