// qboot_b200 host library -- C entry points over the C++ façade for callers that are not C++ (the Python tests and
// bench.py): build the two model problems the reference ships as driver programs and run the whole
// convert -> create_input -> write path on them.  The models are written against the public qboot API only
// (qboot/qboot.hpp) -- the same source compiles against the reference's headers (oracle/ref_capi.cpp holds that copy).
//   model 0: single-correlator 3d Ising  (sample/main.cpp:32-59)
//   model 1: mixed-correlator 3d Ising   (test/src/main.cpp:97-178), `finite` adds the drivers' finite ranges [lb, ub)
//   mode 0: FindContradiction("unit");   mode 1: ExtremalOPE(true, "T", "unit")
#include <array>
#include <chrono>
#include <cstring>
#include <map>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "qboot/qboot.hpp"

using qboot::mp::rational;
using qboot::mp::real;
namespace alg = qboot::algebra;

namespace
{
	using dict = std::map<std::string, rational, std::less<>>;
	using Op = qboot::PrimaryOperator;
	constexpr auto Odd = alg::FunctionSymmetry::Odd;
	constexpr auto Even = alg::FunctionSymmetry::Even;

	qboot::BootstrapEquation model_single(const qboot::Context& c, const dict& d, uint32_t numax, uint32_t maxspin, bool finite)
	{
		Op sig(real(d.at("s")), 0, c), eps(real(d.at("e")), 0, c), unit(c), stress(2, c);
		std::array ssss{sig, sig, sig, sig};
		std::vector<qboot::Sector> secs{{"unit", 1, {real(1)}}, {"e", 1}, {"T", 1}};
		qboot::Sector even("even", 1, qboot::SectorType::Continuous);
		if (finite) even.add_op(0, real("1.409"), real("1.4135"));
		even.add_op(0, real(6));
		even.add_op(2, real(5));
		for (uint32_t l = 4; l <= maxspin; l += 2) even.add_op(l);
		secs.push_back(std::move(even));
		qboot::BootstrapEquation boot(c, std::move(secs), numax);
		qboot::Equation eq(boot, Odd);
		eq.add("unit", unit, ssss);
		eq.add("e", eps, ssss);
		eq.add("T", stress, ssss);
		eq.add("even", ssss);
		boot.add_equation(std::move(eq));
		boot.finish();
		return boot;
	}
	qboot::BootstrapEquation model_mixed(const qboot::Context& c, const dict& d, uint32_t numax, uint32_t maxspin, bool finite)
	{
		Op s(real(d.at("s")), 0, c), e(real(d.at("e")), 0, c), e1(real(d.at("e1")), 0, c), unit(c), stress(2, c);
		std::array ssss{s, s, s, s}, eeee{e, e, e, e}, eses{e, s, e, s}, eess{e, e, s, s}, esse{e, s, s, e};
		std::vector<qboot::Sector> secs{{"unit", 1, {real(1)}}, {"scalar", 2}, {"e1", 2}, {"T", 2, {s.delta(), e.delta()}}};
		qboot::Sector even("even", 2, qboot::SectorType::Continuous);
		if (finite) even.add_op(0, real("3.815"), real("3.845"));
		even.add_op(0, real(6));
		even.add_op(2, real(5));
		for (uint32_t l = 4; l <= maxspin; l += 2) even.add_op(l);
		secs.push_back(even);
		qboot::Sector oddp("odd+", 1, qboot::SectorType::Continuous);
		oddp.add_op(0, real(3));
		if (finite) oddp.add_op(0, real("5.25"), real("5.33"));
		for (uint32_t l = 2; l <= maxspin; l += 2) oddp.add_op(l);
		secs.push_back(oddp);
		qboot::Sector oddm("odd-", 1, qboot::SectorType::Continuous);
		for (uint32_t l = 1; l <= maxspin; l += 2) oddm.add_op(l);
		secs.push_back(oddm);
		qboot::BootstrapEquation boot(c, secs, numax);
		{
			qboot::Equation eq(boot, Odd);
			eq.add("unit", unit, ssss);
			eq.add("scalar", 0, 0, e, ssss);
			eq.add("e1", 0, 0, e1, ssss);
			eq.add("T", 0, 0, stress, ssss);
			eq.add("even", 0, 0, ssss);
			boot.add_equation(eq);
		}
		{
			qboot::Equation eq(boot, Odd);
			eq.add("unit", unit, eeee);
			eq.add("scalar", 1, 1, e, eeee);
			eq.add("e1", 1, 1, e1, eeee);
			eq.add("T", 1, 1, stress, eeee);
			eq.add("even", 1, 1, eeee);
			boot.add_equation(eq);
		}
		{
			qboot::Equation eq(boot, Odd);
			eq.add("scalar", 0, 0, s, eses);
			eq.add("odd+", eses);
			eq.add("odd-", eses);
			boot.add_equation(eq);
		}
		{
			qboot::Equation eq(boot, Odd);
			eq.add("unit", unit, eess);
			eq.add("scalar", 0, 0, s, esse);
			eq.add("scalar", 0, 1, e, eess);
			eq.add("e1", 0, 1, e1, eess);
			eq.add("T", 0, 1, stress, eess);
			eq.add("odd+", esse);
			eq.add("odd-", real(-1), esse);
			eq.add("even", 0, 1, eess);
			boot.add_equation(eq);
		}
		{
			qboot::Equation eq(boot, Even);
			eq.add("unit", unit, eess);
			eq.add("scalar", 0, 0, real(-1), s, esse);
			eq.add("scalar", 0, 1, e, eess);
			eq.add("e1", 0, 1, e1, eess);
			eq.add("T", 0, 1, stress, eess);
			eq.add("odd+", real(-1), esse);
			eq.add("odd-", esse);
			eq.add("even", 0, 1, eess);
			boot.add_equation(eq);
		}
		boot.finish();
		return boot;
	}
	double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
	// QBF_TRACE=1: print the duration of every named phase (the façade's _scoped_event hooks, as the reference's WatchScope does
	// in debug builds) to stderr; per-constraint scopes (names starting with a digit or containing " in ") are summed
	class Trace : public qboot::_event_base
	{
		std::mutex mu_;
		std::map<std::string, double> t0_;
		double small_ = 0;
	public:
		void on_begin(std::string_view tag) override
		{
			std::lock_guard<std::mutex> g(mu_);
			t0_[std::string(tag)] = now();
		}
		void on_end(std::string_view tag) override
		{
			std::lock_guard<std::mutex> g(mu_);
			const std::string k(tag);
			const double dt = now() - t0_[k];
			const bool minor = k.empty() || (k[0] >= '0' && k[0] <= '9') || k.find(" in ") != std::string::npos;
			if (minor)
				small_ += dt;
			else
				std::fprintf(stderr, "[qbf] %-32s %8.1f ms\n", k.c_str(), 1e3 * dt);
		}
		~Trace() override { std::fprintf(stderr, "[qbf] %-32s %8.1f ms (summed over host threads)\n", "per-constraint host tasks", 1e3 * small_); }
	};
	std::unique_ptr<qboot::_event_base> make_trace()
	{
		const char* e = std::getenv("QBF_TRACE");
		if (e && *e && *e != '0') return std::make_unique<Trace>();
		return {};
	}
	thread_local std::string last_error;
}  // namespace

extern "C"
{
	const char* qbf_last_error(void) { return last_error.c_str(); }
	// times[0..2] = seconds of convert / create_input / write; counts[0] = constraints, [1] = N, [2] = free variables
	// outdir may be null / empty: then nothing is written (write phase skipped)
	int qbf_build_pmp(int prec, uint32_t n_max, uint32_t lambda, const char* dim, int model, int mode, int finite, const char* ds,
	                  const char* de, const char* de1, uint32_t numax, uint32_t maxspin, uint32_t threads, const char* outdir,
	                  double* times, int64_t* counts)
	{
		try
		{
			qboot::mp::global_prec = prec;
			qboot::mp::global_rnd = MPFR_RNDN;
			dict d;
			d["s"] = qboot::mp::parse(ds).value();
			d["e"] = qboot::mp::parse(de).value();
			if (de1 && *de1) d["e1"] = qboot::mp::parse(de1).value();
			qboot::Context c(n_max, lambda, rational(std::string_view(dim)), threads);
			auto boot = model == 0 ? model_single(c, d, numax, maxspin, finite != 0) : model_mixed(c, d, numax, maxspin, finite != 0);
			const auto trace = make_trace();
			const double t0 = now();
			auto pmp = mode == 0 ? boot.convert(qboot::FindContradiction("unit"), threads, trace)
			                     : boot.convert(qboot::ExtremalOPE(true, "T", "unit"), threads, trace);
			const double t1 = now();
			if (counts) counts[1] = int64_t(pmp.num_of_variables());
			auto input = std::move(pmp).create_input(threads, trace);
			const double t2 = now();
			if (counts)
			{
				counts[0] = int64_t(input.num_constraints());
				counts[2] = input.num_constraints() ? int64_t(input.constraint(0).num_free()) : 0;
			}
			if (outdir && *outdir) std::move(input).write(qboot::fs::path(outdir), threads, trace);
			const double t3 = now();
			if (times)
			{
				times[0] = t1 - t0;
				times[1] = t2 - t1;
				times[2] = t3 - t2;
			}
		}
		catch (const std::exception& ex)
		{
			last_error = ex.what();
			return -1;
		}
		return 0;
	}
}
