// TEST HARNESS ONLY: C entry point over the façade's ConformalScale (qboot_b200/host/facade_math.cpp) with the output
// layout of oracle/ref_capi.cpp::ref_conformal_scale, so that tests can compare the two bit for bit without a GPU.
#include <cstring>
#include <optional>

#include "qboot/conformal_scale.hpp"

using qboot::mp::real;

extern "C" int qbh_conformal_scale(int prec, uint32_t lambda, int64_t eps_num, int64_t eps_den, uint32_t spin, uint32_t num_poles,
                                   int include_odd, const uint32_t* lb, const uint32_t* ub, uint32_t* xs, uint32_t* scalings,
                                   uint32_t* deltas, uint32_t* bases, int bases_cap)
{
	try
	{
		qb::host::check_prec(prec);
		qboot::mp::global_prec = prec;
		const size_t W = size_t((prec + 31) / 32 + 2);
		auto load = [&](const uint32_t* w) {
			real r;
			std::memcpy(r._words(), w, 4 * W);
			return r;
		};
		auto store = [&](const real& r, uint32_t* w) { std::memcpy(w, r._words(), 4 * W); };
		std::optional<real> end;
		if (ub) end = load(ub);
		const real rho = 3 - qboot::mp::sqrt(8);
		qboot::ConformalScale sc(num_poles, spin, lambda, qboot::mp::rational(eps_num, eps_den), rho, include_odd != 0, load(lb), end);
		const uint32_t D = sc.max_degree();
		auto sp = sc.sample_points();
		auto ss = sc.sample_scalings();
		for (uint32_t k = 0; k <= D; ++k)
		{
			store(sp[k], xs + W * k);
			store(ss[k], scalings + W * k);
			store(sc.get_delta(sp[k]), deltas + W * k);
			if (!(sc.sample_point(k) == sp[k])) return -3;
		}
		auto q = sc.bilinear_bases();
		const uint32_t nb = q.size();
		if (int(nb * nb) > bases_cap) return -2;
		for (uint32_t i = 0; i < nb; ++i)
			for (uint32_t j = 0; j < nb; ++j) store(int32_t(j) <= q[i].degree() ? q[i].at(j) : real{}, bases + W * (size_t(i) * nb + j));
		return int(D);
	}
	catch (...)
	{
		return -1;
	}
}
