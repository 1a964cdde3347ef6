/* gmp.h -- TEST INFRASTRUCTURE ONLY: declarations of the few libgmp entry points the reference's
 * include/marin/engine.h uses (mpz_inits / mpz_clears / mpz_cmp / mpz_set_ui / mpz_import / mpz_export) plus what the
 * adapter test itself calls.  The image ships the libgmp runtime (libgmp.so.10) without its development header; the
 * struct layout and the __gmpz_ symbol names below are GMP's stable ABI (gmp.h of GMP 6.x).  Never used by the product. */
#ifndef MARIN_B200_TEST_GMP_SHIM_H
#define MARIN_B200_TEST_GMP_SHIM_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef unsigned long mp_limb_t;
typedef struct { int _mp_alloc; int _mp_size; mp_limb_t *_mp_d; } __mpz_struct;
typedef __mpz_struct mpz_t[1];
typedef const __mpz_struct *mpz_srcptr;
typedef __mpz_struct *mpz_ptr;
void __gmpz_init(mpz_ptr);
void __gmpz_inits(mpz_ptr, ...);
void __gmpz_clear(mpz_ptr);
void __gmpz_clears(mpz_ptr, ...);
int __gmpz_cmp(mpz_srcptr, mpz_srcptr);
int __gmpz_cmp_ui(mpz_srcptr, unsigned long);
void __gmpz_set_ui(mpz_ptr, unsigned long);
unsigned long __gmpz_get_ui(mpz_srcptr);
void __gmpz_import(mpz_ptr, size_t, int, size_t, int, size_t, const void *);
void *__gmpz_export(void *, size_t *, int, size_t, int, size_t, mpz_srcptr);
void __gmpz_ui_pow_ui(mpz_ptr, unsigned long, unsigned long);
void __gmpz_sub_ui(mpz_ptr, mpz_srcptr, unsigned long);
void __gmpz_add_ui(mpz_ptr, mpz_srcptr, unsigned long);
void __gmpz_mul(mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_mod(mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_mul_2exp(mpz_ptr, mpz_srcptr, unsigned long);
void __gmpz_setbit(mpz_ptr, unsigned long);
void __gmpz_set(mpz_ptr, mpz_srcptr);
void __gmpz_powm_ui(mpz_ptr, mpz_srcptr, unsigned long, mpz_srcptr);
#define mpz_init __gmpz_init
#define mpz_inits __gmpz_inits
#define mpz_clear __gmpz_clear
#define mpz_clears __gmpz_clears
#define mpz_cmp __gmpz_cmp
#define mpz_cmp_ui __gmpz_cmp_ui
#define mpz_set_ui __gmpz_set_ui
#define mpz_get_ui __gmpz_get_ui
#define mpz_import __gmpz_import
#define mpz_export __gmpz_export
#define mpz_ui_pow_ui __gmpz_ui_pow_ui
#define mpz_sub_ui __gmpz_sub_ui
#define mpz_add_ui __gmpz_add_ui
#define mpz_mul __gmpz_mul
#define mpz_mod __gmpz_mod
#define mpz_mul_2exp __gmpz_mul_2exp
#define mpz_setbit __gmpz_setbit
#define mpz_set __gmpz_set
#define mpz_powm_ui __gmpz_powm_ui
#ifdef __cplusplus
}
#endif
#endif
