/* TEST INFRASTRUCTURE ONLY (oracle build) -- not part of the product.
 *
 * ABI shim for the GMP 6.x C interface.  The image carries the runtime library
 * libgmp.so.10 (GMP 6.3.0) but no development headers, so the unmodified
 * reference under /root/reference cannot see <gmp.h>.  This file declares the
 * public structs (documented, ABI-stable layout) and exactly the entry points
 * the reference uses, mapped to the exported `__gmpz_*` / `__gmpq_*` symbols.
 * Written from the GMP manual's prototypes; contains no GMP source.
 */
#ifndef QB_ORACLE_SHIM_GMP_H_
#define QB_ORACLE_SHIM_GMP_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned long mp_limb_t;
typedef long mp_limb_signed_t;
typedef unsigned long mp_bitcnt_t;
typedef long mp_size_t;
typedef long mp_exp_t;

typedef struct
{
	int _mp_alloc;
	int _mp_size;
	mp_limb_t* _mp_d;
} __mpz_struct;
typedef __mpz_struct mpz_t[1];
typedef __mpz_struct* mpz_ptr;
typedef const __mpz_struct* mpz_srcptr;

typedef struct
{
	__mpz_struct _mp_num;
	__mpz_struct _mp_den;
} __mpq_struct;
typedef __mpq_struct mpq_t[1];
typedef __mpq_struct* mpq_ptr;
typedef const __mpq_struct* mpq_srcptr;

#define __GNU_MP_VERSION 6
#define __GNU_MP_VERSION_MINOR 3
#define __GNU_MP_VERSION_PATCHLEVEL 0

/* ---- mpz ---- */
#define QB_Z(name) __gmpz_##name
#define mpz_init QB_Z(init)
#define mpz_clear QB_Z(clear)
#define mpz_init_set QB_Z(init_set)
#define mpz_init_set_d QB_Z(init_set_d)
#define mpz_init_set_si QB_Z(init_set_si)
#define mpz_init_set_ui QB_Z(init_set_ui)
#define mpz_init_set_str QB_Z(init_set_str)
#define mpz_set QB_Z(set)
#define mpz_set_d QB_Z(set_d)
#define mpz_set_si QB_Z(set_si)
#define mpz_set_ui QB_Z(set_ui)
#define mpz_set_str QB_Z(set_str)
#define mpz_swap QB_Z(swap)
#define mpz_abs QB_Z(abs)
#define mpz_neg QB_Z(neg)
#define mpz_add QB_Z(add)
#define mpz_add_ui QB_Z(add_ui)
#define mpz_sub QB_Z(sub)
#define mpz_sub_ui QB_Z(sub_ui)
#define mpz_ui_sub QB_Z(ui_sub)
#define mpz_mul QB_Z(mul)
#define mpz_mul_si QB_Z(mul_si)
#define mpz_mul_ui QB_Z(mul_ui)
#define mpz_mul_2exp QB_Z(mul_2exp)
#define mpz_addmul QB_Z(addmul)
#define mpz_addmul_ui QB_Z(addmul_ui)
#define mpz_submul QB_Z(submul)
#define mpz_submul_ui QB_Z(submul_ui)
#define mpz_cdiv_q QB_Z(cdiv_q)
#define mpz_fdiv_q QB_Z(fdiv_q)
#define mpz_fdiv_q_ui QB_Z(fdiv_q_ui)
#define mpz_fdiv_r QB_Z(fdiv_r)
#define mpz_fdiv_r_ui QB_Z(fdiv_r_ui)
#define mpz_fdiv_ui QB_Z(fdiv_ui)
#define mpz_tdiv_q QB_Z(tdiv_q)
#define mpz_cmp QB_Z(cmp)
#define mpz_cmp_d QB_Z(cmp_d)
#define mpz_cmp_si QB_Z(cmp_si)
#define mpz_cmp_ui QB_Z(cmp_ui)
#define mpz_cmpabs QB_Z(cmpabs)
#define mpz_cmpabs_ui QB_Z(cmpabs_ui)
#define mpz_divisible_p QB_Z(divisible_p)
#define mpz_divisible_ui_p QB_Z(divisible_ui_p)
#define mpz_fac_ui QB_Z(fac_ui)
#define mpz_gcd QB_Z(gcd)
#define mpz_lcm QB_Z(lcm)
#define mpz_get_d QB_Z(get_d)
#define mpz_get_si QB_Z(get_si)
#define mpz_get_ui QB_Z(get_ui)
#define mpz_get_str QB_Z(get_str)
#define mpz_pow_ui QB_Z(pow_ui)
#define mpz_ui_pow_ui QB_Z(ui_pow_ui)
#define mpz_sizeinbase QB_Z(sizeinbase)

void mpz_init(mpz_ptr);
void mpz_clear(mpz_ptr);
void mpz_init_set(mpz_ptr, mpz_srcptr);
void mpz_init_set_d(mpz_ptr, double);
void mpz_init_set_si(mpz_ptr, long);
void mpz_init_set_ui(mpz_ptr, unsigned long);
int mpz_init_set_str(mpz_ptr, const char*, int);
void mpz_set(mpz_ptr, mpz_srcptr);
void mpz_set_d(mpz_ptr, double);
void mpz_set_si(mpz_ptr, long);
void mpz_set_ui(mpz_ptr, unsigned long);
int mpz_set_str(mpz_ptr, const char*, int);
void mpz_swap(mpz_ptr, mpz_ptr);
void mpz_abs(mpz_ptr, mpz_srcptr);
void mpz_neg(mpz_ptr, mpz_srcptr);
void mpz_add(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_add_ui(mpz_ptr, mpz_srcptr, unsigned long);
void mpz_sub(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_sub_ui(mpz_ptr, mpz_srcptr, unsigned long);
void mpz_ui_sub(mpz_ptr, unsigned long, mpz_srcptr);
void mpz_mul(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_mul_si(mpz_ptr, mpz_srcptr, long);
void mpz_mul_ui(mpz_ptr, mpz_srcptr, unsigned long);
void mpz_mul_2exp(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
void mpz_addmul(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_addmul_ui(mpz_ptr, mpz_srcptr, unsigned long);
void mpz_submul(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_submul_ui(mpz_ptr, mpz_srcptr, unsigned long);
void mpz_cdiv_q(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_fdiv_q(mpz_ptr, mpz_srcptr, mpz_srcptr);
unsigned long mpz_fdiv_q_ui(mpz_ptr, mpz_srcptr, unsigned long);
void mpz_fdiv_r(mpz_ptr, mpz_srcptr, mpz_srcptr);
unsigned long mpz_fdiv_r_ui(mpz_ptr, mpz_srcptr, unsigned long);
unsigned long mpz_fdiv_ui(mpz_srcptr, unsigned long);
void mpz_tdiv_q(mpz_ptr, mpz_srcptr, mpz_srcptr);
int mpz_cmp(mpz_srcptr, mpz_srcptr);
int mpz_cmp_d(mpz_srcptr, double);
int mpz_cmp_si(mpz_srcptr, long);
int mpz_cmp_ui(mpz_srcptr, unsigned long);
int mpz_cmpabs(mpz_srcptr, mpz_srcptr);
int mpz_cmpabs_ui(mpz_srcptr, unsigned long);
int mpz_divisible_p(mpz_srcptr, mpz_srcptr);
int mpz_divisible_ui_p(mpz_srcptr, unsigned long);
void mpz_fac_ui(mpz_ptr, unsigned long);
void mpz_gcd(mpz_ptr, mpz_srcptr, mpz_srcptr);
void mpz_lcm(mpz_ptr, mpz_srcptr, mpz_srcptr);
double mpz_get_d(mpz_srcptr);
long mpz_get_si(mpz_srcptr);
unsigned long mpz_get_ui(mpz_srcptr);
char* mpz_get_str(char*, int, mpz_srcptr);
void mpz_pow_ui(mpz_ptr, mpz_srcptr, unsigned long);
void mpz_ui_pow_ui(mpz_ptr, unsigned long, unsigned long);
size_t mpz_sizeinbase(mpz_srcptr, int);

#define mpz_sgn(Z) ((Z)->_mp_size < 0 ? -1 : (Z)->_mp_size > 0)
#define mpz_odd_p(z) (((z)->_mp_size != 0) & (int)((z)->_mp_d[0]))
#define mpz_even_p(z) (!mpz_odd_p(z))

/* ---- mpq ---- */
#define QB_Q(name) __gmpq_##name
#define mpq_init QB_Q(init)
#define mpq_clear QB_Q(clear)
#define mpq_set QB_Q(set)
#define mpq_set_d QB_Q(set_d)
#define mpq_set_si QB_Q(set_si)
#define mpq_set_ui QB_Q(set_ui)
#define mpq_set_str QB_Q(set_str)
#define mpq_set_z QB_Q(set_z)
#define mpq_swap QB_Q(swap)
#define mpq_abs QB_Q(abs)
#define mpq_neg QB_Q(neg)
#define mpq_add QB_Q(add)
#define mpq_sub QB_Q(sub)
#define mpq_mul QB_Q(mul)
#define mpq_div QB_Q(div)
#define mpq_inv QB_Q(inv)
#define mpq_canonicalize QB_Q(canonicalize)
#define mpq_cmp QB_Q(cmp)
#define mpq_cmp_si QB_Q(cmp_si)
#define mpq_cmp_ui QB_Q(cmp_ui)
#define mpq_cmp_z QB_Q(cmp_z)
#define mpq_get_str QB_Q(get_str)

void mpq_init(mpq_ptr);
void mpq_clear(mpq_ptr);
void mpq_set(mpq_ptr, mpq_srcptr);
void mpq_set_d(mpq_ptr, double);
void mpq_set_si(mpq_ptr, long, unsigned long);
void mpq_set_ui(mpq_ptr, unsigned long, unsigned long);
int mpq_set_str(mpq_ptr, const char*, int);
void mpq_set_z(mpq_ptr, mpz_srcptr);
void mpq_swap(mpq_ptr, mpq_ptr);
void mpq_abs(mpq_ptr, mpq_srcptr);
void mpq_neg(mpq_ptr, mpq_srcptr);
void mpq_add(mpq_ptr, mpq_srcptr, mpq_srcptr);
void mpq_sub(mpq_ptr, mpq_srcptr, mpq_srcptr);
void mpq_mul(mpq_ptr, mpq_srcptr, mpq_srcptr);
void mpq_div(mpq_ptr, mpq_srcptr, mpq_srcptr);
void mpq_inv(mpq_ptr, mpq_srcptr);
void mpq_canonicalize(mpq_ptr);
int mpq_cmp(mpq_srcptr, mpq_srcptr);
int mpq_cmp_si(mpq_srcptr, long, unsigned long);
int mpq_cmp_ui(mpq_srcptr, unsigned long, unsigned long);
int mpq_cmp_z(mpq_srcptr, mpz_srcptr);
char* mpq_get_str(char*, int, mpq_srcptr);

#define mpq_numref(Q) (&((Q)->_mp_num))
#define mpq_denref(Q) (&((Q)->_mp_den))
#define mpq_sgn(Q) ((Q)->_mp_num._mp_size < 0 ? -1 : (Q)->_mp_num._mp_size > 0)

#ifdef __cplusplus
}
#endif

#endif /* QB_ORACLE_SHIM_GMP_H_ */
