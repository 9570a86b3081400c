// qboot_b200 façade -- host arithmetic on number records (implemented in qboot_b200/host/hostops.cpp).
//
// A number is the record of include/qboot_b200.h: W = L + 2 uint32 words, L = ceil(prec / 32) (kind|sign, exponent,
// limbs).  Every function returns the exact result rounded once to nearest-even at `prec` bits -- the contract of the
// mpfr_* calls behind the reference's mp::real (include/qboot/mp/real.hpp:324-361,688-804) -- and the basic operations
// are the very `__host__ __device__` routines the CUDA kernels run (qboot_b200/csrc/mpf.cuh) compiled for the host, so
// host-side and device-side values are interchangeable bit for bit.  Output may alias an input.
#ifndef QBOOT_B200_HOSTOPS_HPP_
#define QBOOT_B200_HOSTOPS_HPP_

#include <cstdint>
#include <string>
#include <string_view>

namespace qb::host
{
	constexpr int kMaxWords = 50;  // 1536 bits
	// limb count for a precision, or 0 if no kernels were compiled for it (qb_precision_supported)
	int limbs_for(int prec);
	void check_prec(int prec);  // throws std::invalid_argument for an unsupported precision

	void h_add(uint32_t* r, const uint32_t* a, const uint32_t* b, int prec);
	void h_sub(uint32_t* r, const uint32_t* a, const uint32_t* b, int prec);
	void h_mul(uint32_t* r, const uint32_t* a, const uint32_t* b, int prec);
	void h_div(uint32_t* r, const uint32_t* a, const uint32_t* b, int prec);
	void h_fma(uint32_t* r, const uint32_t* a, const uint32_t* b, const uint32_t* c, int prec);  // a * b + c
	void h_fms(uint32_t* r, const uint32_t* a, const uint32_t* b, const uint32_t* c, int prec);  // a * b - c
	void h_sqrt(uint32_t* r, const uint32_t* a, int prec);
	void h_neg(uint32_t* r, const uint32_t* a, int prec);
	void h_abs(uint32_t* r, const uint32_t* a, int prec);
	void h_set_si(uint32_t* r, int64_t v, int prec);
	void h_set_ui(uint32_t* r, uint64_t v, int prec);
	void h_set_d(uint32_t* r, double v, int prec);
	void h_set_si_2exp(uint32_t* r, int64_t v, int64_t e, int prec);  // v * 2^e
	void h_set_q(uint32_t* r, int64_t num, int64_t den, int prec);
	void h_mul_si(uint32_t* r, const uint32_t* a, int64_t v, int prec);
	void h_div_si(uint32_t* r, const uint32_t* a, int64_t v, int prec);
	void h_add_si(uint32_t* r, const uint32_t* a, int64_t v, int prec);
	void h_si_sub(uint32_t* r, int64_t v, const uint32_t* a, int prec);  // v - a
	void h_si_div(uint32_t* r, int64_t v, const uint32_t* a, int prec);  // v / a
	void h_mul_q(uint32_t* r, const uint32_t* a, int64_t num, int64_t den, int prec);
	void h_div_q(uint32_t* r, const uint32_t* a, int64_t num, int64_t den, int prec);
	void h_add_q(uint32_t* r, const uint32_t* a, int64_t num, int64_t den, int prec);
	void h_sub_q(uint32_t* r, const uint32_t* a, int64_t num, int64_t den, int prec);
	void h_mul_2si(uint32_t* r, const uint32_t* a, int64_t e, int prec);
	int h_cmp(const uint32_t* a, const uint32_t* b, int prec);      // -1, 0, 1 (nan compares as 0 like mpfr_cmp)
	int h_cmpabs(const uint32_t* a, const uint32_t* b, int prec);
	int h_cmp_si(const uint32_t* a, int64_t v, int prec);
	int h_cmp_q(const uint32_t* a, int64_t num, int64_t den, int prec);
	bool h_is_integer(const uint32_t* a, int prec);
	void h_floor(uint32_t* r, const uint32_t* a, int prec);
	void h_ceil(uint32_t* r, const uint32_t* a, int prec);
	void h_round(uint32_t* r, const uint32_t* a, int prec);
	double h_get_d(const uint32_t* a, int prec);
	int64_t h_get_si(const uint32_t* a, int prec);  // truncated toward zero
	// correctly rounded transcendentals
	void h_exp(uint32_t* r, const uint32_t* a, int prec);
	void h_log(uint32_t* r, const uint32_t* a, int prec);
	void h_pow(uint32_t* r, const uint32_t* a, const uint32_t* b, int prec);
	void h_pow_si(uint32_t* r, const uint32_t* a, int64_t n, int prec);
	void h_ui_pow(uint32_t* r, uint64_t base, const uint32_t* b, int prec);
	void h_const_pi(uint32_t* r, int prec);
	void h_const_log2(uint32_t* r, int prec);
	void h_gamma_inc(uint32_t* r, const uint32_t* a, const uint32_t* x, int prec);  // Gamma(a, x); a = 0 only (others throw)
	// decimal / "0x" text -> number (mpfr_set_str base 0 semantics for plain decimals); false on a malformed string
	bool h_from_string(uint32_t* r, std::string_view s, int prec);
	// the `ndigits` leading decimal digits of |a| rounded to nearest (ties to even) and the decimal exponent e with
	// |a| = 0.d1 d2 ... * 10^e   (mpfr_get_str(0, &e, 10, ndigits, a, MPFR_RNDN)); a must be regular
	std::string h_get_digits(const uint32_t* a, int prec, size_t ndigits, int64_t* e);
}  // namespace qb::host

#endif  // QBOOT_B200_HOSTOPS_HPP_
