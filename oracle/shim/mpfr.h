/* TEST INFRASTRUCTURE ONLY (oracle build) -- not part of the product.
 *
 * ABI shim for the MPFR 4.x C interface.  The image carries the runtime
 * library libmpfr.so.6 (MPFR 4.2.1) but no development headers.  This file
 * declares the public struct (documented layout) and exactly the functions the
 * unmodified reference under /root/reference calls, plus a few used by the
 * oracle's own C-ABI driver.  Written from the MPFR manual's prototypes;
 * contains no MPFR source.
 */
#ifndef QB_ORACLE_SHIM_MPFR_H_
#define QB_ORACLE_SHIM_MPFR_H_

#include <stdint.h>

#include "gmp.h"

#ifdef __cplusplus
extern "C" {
#endif

#define MPFR_VERSION_MAJOR 4
#define MPFR_VERSION_MINOR 2
#define MPFR_VERSION_PATCHLEVEL 1

typedef long mpfr_prec_t;
typedef unsigned long mpfr_uprec_t;
typedef int mpfr_sign_t;
typedef long mpfr_exp_t;
typedef unsigned long mpfr_uexp_t;

typedef enum
{
	MPFR_RNDN = 0,
	MPFR_RNDZ,
	MPFR_RNDU,
	MPFR_RNDD,
	MPFR_RNDA,
	MPFR_RNDF,
	MPFR_RNDNA = -1
} mpfr_rnd_t;

typedef struct
{
	mpfr_prec_t _mpfr_prec;
	mpfr_sign_t _mpfr_sign;
	mpfr_exp_t _mpfr_exp;
	mp_limb_t* _mpfr_d;
} __mpfr_struct;
typedef __mpfr_struct mpfr_t[1];
typedef __mpfr_struct* mpfr_ptr;
typedef const __mpfr_struct* mpfr_srcptr;

/* intmax_t variants are exported under mangled names */
#define mpfr_set_sj __gmpfr_set_sj
#define mpfr_set_sj_2exp __gmpfr_set_sj_2exp
#define mpfr_set_uj __gmpfr_set_uj
#define mpfr_set_uj_2exp __gmpfr_set_uj_2exp
#define mpfr_get_sj __gmpfr_mpfr_get_sj
#define mpfr_get_uj __gmpfr_mpfr_get_uj

void mpfr_init2(mpfr_ptr, mpfr_prec_t);
void mpfr_clear(mpfr_ptr);
void mpfr_swap(mpfr_ptr, mpfr_ptr);
mpfr_prec_t mpfr_get_prec(mpfr_srcptr);
mpfr_exp_t mpfr_get_exp(mpfr_srcptr);

int mpfr_set(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_set_d(mpfr_ptr, double, mpfr_rnd_t);
int mpfr_set_q(mpfr_ptr, mpq_srcptr, mpfr_rnd_t);
int mpfr_set_z(mpfr_ptr, mpz_srcptr, mpfr_rnd_t);
int mpfr_set_z_2exp(mpfr_ptr, mpz_srcptr, mpfr_exp_t, mpfr_rnd_t);
int mpfr_set_si(mpfr_ptr, long, mpfr_rnd_t);
int mpfr_set_ui(mpfr_ptr, unsigned long, mpfr_rnd_t);
int mpfr_set_si_2exp(mpfr_ptr, long, mpfr_exp_t, mpfr_rnd_t);
int mpfr_set_ui_2exp(mpfr_ptr, unsigned long, mpfr_exp_t, mpfr_rnd_t);
int mpfr_set_sj(mpfr_ptr, intmax_t, mpfr_rnd_t);
int mpfr_set_uj(mpfr_ptr, uintmax_t, mpfr_rnd_t);
int mpfr_set_sj_2exp(mpfr_ptr, intmax_t, intmax_t, mpfr_rnd_t);
int mpfr_set_uj_2exp(mpfr_ptr, uintmax_t, intmax_t, mpfr_rnd_t);
int mpfr_set_str(mpfr_ptr, const char*, int, mpfr_rnd_t);
void mpfr_set_zero(mpfr_ptr, int);
void mpfr_set_inf(mpfr_ptr, int);
void mpfr_set_nan(mpfr_ptr);

double mpfr_get_d(mpfr_srcptr, mpfr_rnd_t);
long mpfr_get_si(mpfr_srcptr, mpfr_rnd_t);
unsigned long mpfr_get_ui(mpfr_srcptr, mpfr_rnd_t);
intmax_t mpfr_get_sj(mpfr_srcptr, mpfr_rnd_t);
uintmax_t mpfr_get_uj(mpfr_srcptr, mpfr_rnd_t);
int mpfr_get_z(mpz_ptr, mpfr_srcptr, mpfr_rnd_t);
mpfr_exp_t mpfr_get_z_2exp(mpz_ptr, mpfr_srcptr);
void mpfr_get_q(mpq_ptr, mpfr_srcptr);
char* mpfr_get_str(char*, mpfr_exp_t*, int, size_t, mpfr_srcptr, mpfr_rnd_t);
void mpfr_free_str(char*);

int mpfr_nan_p(mpfr_srcptr);
int mpfr_inf_p(mpfr_srcptr);
int mpfr_zero_p(mpfr_srcptr);
int mpfr_integer_p(mpfr_srcptr);
int mpfr_sgn(mpfr_srcptr);
int mpfr_signbit(mpfr_srcptr);

int mpfr_cmp(mpfr_srcptr, mpfr_srcptr);
int mpfr_cmpabs(mpfr_srcptr, mpfr_srcptr);
int mpfr_cmp_d(mpfr_srcptr, double);
int mpfr_cmp_q(mpfr_srcptr, mpq_srcptr);
int mpfr_cmp_z(mpfr_srcptr, mpz_srcptr);
int mpfr_cmp_si(mpfr_srcptr, long);
int mpfr_cmp_ui(mpfr_srcptr, unsigned long);

int mpfr_abs(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_neg(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);

int mpfr_add(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_add_d(mpfr_ptr, mpfr_srcptr, double, mpfr_rnd_t);
int mpfr_add_q(mpfr_ptr, mpfr_srcptr, mpq_srcptr, mpfr_rnd_t);
int mpfr_add_z(mpfr_ptr, mpfr_srcptr, mpz_srcptr, mpfr_rnd_t);
int mpfr_add_si(mpfr_ptr, mpfr_srcptr, long, mpfr_rnd_t);
int mpfr_add_ui(mpfr_ptr, mpfr_srcptr, unsigned long, mpfr_rnd_t);
int mpfr_sub(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_sub_d(mpfr_ptr, mpfr_srcptr, double, mpfr_rnd_t);
int mpfr_sub_q(mpfr_ptr, mpfr_srcptr, mpq_srcptr, mpfr_rnd_t);
int mpfr_sub_z(mpfr_ptr, mpfr_srcptr, mpz_srcptr, mpfr_rnd_t);
int mpfr_sub_si(mpfr_ptr, mpfr_srcptr, long, mpfr_rnd_t);
int mpfr_sub_ui(mpfr_ptr, mpfr_srcptr, unsigned long, mpfr_rnd_t);
int mpfr_d_sub(mpfr_ptr, double, mpfr_srcptr, mpfr_rnd_t);
int mpfr_z_sub(mpfr_ptr, mpz_srcptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_si_sub(mpfr_ptr, long, mpfr_srcptr, mpfr_rnd_t);
int mpfr_ui_sub(mpfr_ptr, unsigned long, mpfr_srcptr, mpfr_rnd_t);
int mpfr_mul(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_mul_d(mpfr_ptr, mpfr_srcptr, double, mpfr_rnd_t);
int mpfr_mul_q(mpfr_ptr, mpfr_srcptr, mpq_srcptr, mpfr_rnd_t);
int mpfr_mul_z(mpfr_ptr, mpfr_srcptr, mpz_srcptr, mpfr_rnd_t);
int mpfr_mul_si(mpfr_ptr, mpfr_srcptr, long, mpfr_rnd_t);
int mpfr_mul_ui(mpfr_ptr, mpfr_srcptr, unsigned long, mpfr_rnd_t);
int mpfr_mul_2si(mpfr_ptr, mpfr_srcptr, long, mpfr_rnd_t);
int mpfr_div(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_div_d(mpfr_ptr, mpfr_srcptr, double, mpfr_rnd_t);
int mpfr_div_q(mpfr_ptr, mpfr_srcptr, mpq_srcptr, mpfr_rnd_t);
int mpfr_div_z(mpfr_ptr, mpfr_srcptr, mpz_srcptr, mpfr_rnd_t);
int mpfr_div_si(mpfr_ptr, mpfr_srcptr, long, mpfr_rnd_t);
int mpfr_div_ui(mpfr_ptr, mpfr_srcptr, unsigned long, mpfr_rnd_t);
int mpfr_d_div(mpfr_ptr, double, mpfr_srcptr, mpfr_rnd_t);
int mpfr_si_div(mpfr_ptr, long, mpfr_srcptr, mpfr_rnd_t);
int mpfr_ui_div(mpfr_ptr, unsigned long, mpfr_srcptr, mpfr_rnd_t);

int mpfr_fma(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_fms(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_fmma(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_fmms(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);

int mpfr_sqrt(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_sqrt_ui(mpfr_ptr, unsigned long, mpfr_rnd_t);
int mpfr_pow(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_pow_si(mpfr_ptr, mpfr_srcptr, long, mpfr_rnd_t);
int mpfr_pow_ui(mpfr_ptr, mpfr_srcptr, unsigned long, mpfr_rnd_t);
int mpfr_ui_pow(mpfr_ptr, unsigned long, mpfr_srcptr, mpfr_rnd_t);
int mpfr_exp(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_log(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_sin(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_cos(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_tan(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_sec(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_csc(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_cot(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_asin(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_acos(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_atan(mpfr_ptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_atan2(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_gamma_inc(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_const_pi(mpfr_ptr, mpfr_rnd_t);
int mpfr_const_log2(mpfr_ptr, mpfr_rnd_t);

int mpfr_fmod(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_dim(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_min(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_max(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
int mpfr_round(mpfr_ptr, mpfr_srcptr);
int mpfr_floor(mpfr_ptr, mpfr_srcptr);
int mpfr_ceil(mpfr_ptr, mpfr_srcptr);
void mpfr_nextabove(mpfr_ptr);
void mpfr_nextbelow(mpfr_ptr);
void mpfr_nexttoward(mpfr_ptr, mpfr_srcptr);

#ifdef __cplusplus
}
#endif

#endif /* QB_ORACLE_SHIM_MPFR_H_ */
