// TEST HARNESS ONLY: C entry points over the façade's host arithmetic (qboot_b200/host/hostops.cpp) so that the CPU
// test-suite can pin every operation against MPFR through oracle/_ref, with the op codes of oracle/ref_capi.cpp.
#include <cstdint>
#include <cstring>
#include <string>

#include "qboot/b200/hostops.hpp"

using namespace qb::host;

extern "C" int qbh_hostop(int prec, int op, int n, const uint32_t* a, const uint32_t* b, const uint32_t* c, const int64_t* ia,
                          const int64_t* ib, uint32_t* out)
{
	try
	{
		check_prec(prec);
		const size_t W = size_t((prec + 31) / 32 + 2);
		uint32_t zero[kMaxWords] = {0};
		for (int i = 0; i < n; ++i)
		{
			const uint32_t* x = a ? a + W * size_t(i) : zero;
			const uint32_t* y = b ? b + W * size_t(i) : zero;
			const uint32_t* z = c ? c + W * size_t(i) : zero;
			int64_t p = ia ? ia[i] : 0, q = ib ? ib[i] : 1;
			uint32_t r[kMaxWords] = {0};
			switch (op)
			{
			case 0: h_add(r, x, y, prec); break;
			case 1: h_sub(r, x, y, prec); break;
			case 2: h_mul(r, x, y, prec); break;
			case 3: h_div(r, x, y, prec); break;
			case 4: h_fma(r, x, y, z, prec); break;
			case 5: h_sqrt(r, x, prec); break;
			case 6: h_mul_si(r, x, p, prec); break;
			case 7: h_add_si(r, x, p, prec); break;
			case 8: h_div_si(r, x, p, prec); break;
			case 9: h_mul_q(r, x, p, q, prec); break;
			case 10: h_div_q(r, x, p, q, prec); break;
			case 11: h_add_q(r, x, p, q, prec); break;
			case 12: h_set_q(r, p, q, prec); break;
			case 13: h_pow(r, x, y, prec); break;
			case 14: h_ui_pow(r, uint64_t(p), x, prec); break;
			case 15: h_exp(r, x, prec); break;
			case 16: h_log(r, x, prec); break;
			case 17: h_gamma_inc(r, x, y, prec); break;
			case 18: h_si_sub(r, p, x, prec); break;
			case 19: h_si_div(r, p, x, prec); break;
			case 20: h_sub_q(r, x, p, q, prec); break;
			case 21: h_const_pi(r, prec); break;
			case 22: h_set_si_2exp(r, p, q, prec); break;
			case 23: r[0] = uint32_t(h_cmp(x, y, prec) + 1); break;
			case 24: r[0] = h_is_integer(x, prec) ? 1u : 0u; break;
			case 25: h_floor(r, x, prec); break;
			case 26: h_pow_si(r, x, p, prec); break;
			default: return -1;
			}
			std::memcpy(out + W * size_t(i), r, 4 * W);
		}
	}
	catch (...)
	{
		return -2;
	}
	return 0;
}
extern "C" int qbh_from_str(int prec, const char* s, uint32_t* out)
{
	try
	{
		return h_from_string(out, s, prec) ? 0 : -1;
	}
	catch (...)
	{
		return -2;
	}
}
// digits <= 0: the SDPB writer's 3 + int(prec * 0.302); text in the reference's default-float format
extern "C" int qbh_digits(int prec, const uint32_t* x, int ndigits, char* buf, int buflen, int64_t* e10)
{
	try
	{
		std::string s = h_get_digits(x, prec, size_t(ndigits), e10);
		if (int(s.size()) + 1 > buflen) return -1;
		std::memcpy(buf, s.c_str(), s.size() + 1);
		return int(s.size());
	}
	catch (...)
	{
		return -2;
	}
}
#include <sstream>

#include "qboot/mp/real.hpp"
// the reference's ostream text at `digits` significant digits (digits <= 0: the SDPB writer's 3 + int(prec * 0.302))
extern "C" int qbh_to_str(int prec, const uint32_t* x, int digits, char* buf, int buflen)
{
	try
	{
		qboot::mp::global_prec = prec;
		qboot::mp::real r;
		std::memcpy(r._words(), x, 4 * size_t((prec + 31) / 32 + 2));
		std::ostringstream os;
		os.precision(digits > 0 ? digits : 3 + int(double(prec) * 0.302));
		os << r;
		auto s = os.str();
		if (int(s.size()) + 1 > buflen) return -1;
		std::memcpy(buf, s.c_str(), s.size() + 1);
		return int(s.size());
	}
	catch (...)
	{
		return -2;
	}
}
