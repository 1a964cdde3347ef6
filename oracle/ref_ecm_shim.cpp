// ref_ecm_shim.cpp -- TEST INFRASTRUCTURE ONLY.
// Pins the ECM driver's seed -> sigma -> twisted-Edwards curve chain to the REFERENCE'S OWN CODE: the Makefile cuts
// three line ranges out of /root/reference/src/modes/RunEcmTwistedEdwards.cpp where it lies (seed mixing :857-887,
// modular helpers :2082-2119, the Suyama-sigma curve block :2437-2586) into oracle/_ref/*.inc, this file #includes
// them inside a scaffold that supplies the few names they refer to, and the .inc files are deleted again after the
// compile: only oracle/_ref/libecm_ref.so remains (git-ignored).  No reference source is copied into the repository.
// GMP: the image has the libgmp runtime but no headers, so the handful of mpz_* entry points and a minimal mpz_class
// (the operators those line ranges use) are declared here.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>

extern "C" {
typedef struct { int _mp_alloc; int _mp_size; void *_mp_d; } __mpz_struct;
typedef __mpz_struct mpz_t[1];
void __gmpz_init(__mpz_struct *); void __gmpz_clear(__mpz_struct *);
void __gmpz_set(__mpz_struct *, const __mpz_struct *); void __gmpz_set_ui(__mpz_struct *, unsigned long); void __gmpz_set_si(__mpz_struct *, long);
int __gmpz_set_str(__mpz_struct *, const char *, int); char *__gmpz_get_str(char *, int, const __mpz_struct *);
void __gmpz_add(__mpz_struct *, const __mpz_struct *, const __mpz_struct *); void __gmpz_sub(__mpz_struct *, const __mpz_struct *, const __mpz_struct *);
void __gmpz_mul(__mpz_struct *, const __mpz_struct *, const __mpz_struct *); void __gmpz_mod(__mpz_struct *, const __mpz_struct *, const __mpz_struct *);
void __gmpz_add_ui(__mpz_struct *, const __mpz_struct *, unsigned long); void __gmpz_mul_2exp(__mpz_struct *, const __mpz_struct *, unsigned long);
void __gmpz_tdiv_r(__mpz_struct *, const __mpz_struct *, const __mpz_struct *);
int __gmpz_invert(__mpz_struct *, const __mpz_struct *, const __mpz_struct *); void __gmpz_gcd(__mpz_struct *, const __mpz_struct *, const __mpz_struct *);
int __gmpz_cmp(const __mpz_struct *, const __mpz_struct *); int __gmpz_cmp_si(const __mpz_struct *, long);
}
#define mpz_mod __gmpz_mod
#define mpz_mul __gmpz_mul
#define mpz_invert __gmpz_invert
#define mpz_gcd __gmpz_gcd
#define mpz_set_str __gmpz_set_str
#define mpz_sgn(z) ((z)->_mp_size < 0 ? -1 : (z)->_mp_size > 0)

class mpz_class
{
	mpz_t z;
public:
	mpz_class() { __gmpz_init(z); }
	mpz_class(int v) { __gmpz_init(z); __gmpz_set_si(z, v); }
	mpz_class(long v) { __gmpz_init(z); __gmpz_set_si(z, v); }
	mpz_class(unsigned long v) { __gmpz_init(z); __gmpz_set_ui(z, v); }
	mpz_class(const mpz_class &o) { __gmpz_init(z); __gmpz_set(z, o.z); }
	~mpz_class() { __gmpz_clear(z); }
	mpz_class &operator=(const mpz_class &o) { __gmpz_set(z, o.z); return *this; }
	__mpz_struct *get_mpz_t() { return z; }
	const __mpz_struct *get_mpz_t() const { return z; }
	std::string get_str(int base) const { char *s = __gmpz_get_str(nullptr, base, z); std::string r(s); free(s); return r; }
	mpz_class &operator<<=(unsigned long k) { __gmpz_mul_2exp(z, z, k); return *this; }
	mpz_class &operator+=(unsigned long v) { __gmpz_add_ui(z, z, v); return *this; }
	mpz_class &operator+=(int v) { if (v >= 0) __gmpz_add_ui(z, z, (unsigned long)v); else { mpz_class t(v); __gmpz_add(z, z, t.z); } return *this; }
	mpz_class &operator+=(const mpz_class &o) { __gmpz_add(z, z, o.z); return *this; }
	mpz_class &operator%=(const mpz_class &o) { __gmpz_tdiv_r(z, z, o.z); return *this; }   // truncating, like gmpxx
	friend mpz_class operator+(const mpz_class &a, const mpz_class &b) { mpz_class r; __gmpz_add(r.z, a.z, b.z); return r; }
	friend mpz_class operator-(const mpz_class &a, const mpz_class &b) { mpz_class r; __gmpz_sub(r.z, a.z, b.z); return r; }
	friend mpz_class operator<<(const mpz_class &a, unsigned long k) { mpz_class r; __gmpz_mul_2exp(r.z, a.z, k); return r; }
	friend bool operator<(const mpz_class &a, const mpz_class &b) { return __gmpz_cmp(a.z, b.z) < 0; }
	friend bool operator>(const mpz_class &a, const mpz_class &b) { return __gmpz_cmp(a.z, b.z) > 0; }
	friend bool operator<=(const mpz_class &a, const mpz_class &b) { return __gmpz_cmp(a.z, b.z) <= 0; }
	friend bool operator==(const mpz_class &a, const mpz_class &b) { return __gmpz_cmp(a.z, b.z) == 0; }
	friend bool operator!=(const mpz_class &a, const mpz_class &b) { return __gmpz_cmp(a.z, b.z) != 0; }
	friend bool operator<(const mpz_class &a, int b) { return __gmpz_cmp_si(a.z, b) < 0; }
	friend bool operator>(const mpz_class &a, int b) { return __gmpz_cmp_si(a.z, b) > 0; }
	friend bool operator<=(const mpz_class &a, int b) { return __gmpz_cmp_si(a.z, b) <= 0; }
	friend bool operator==(const mpz_class &a, int b) { return __gmpz_cmp_si(a.z, b) == 0; }
	friend bool operator!=(const mpz_class &a, int b) { return __gmpz_cmp_si(a.z, b) != 0; }
};

using std::string;
namespace { struct engine_stub {}; struct options_stub { std::string sigma, sigma_hex; uint32_t curves_tested_for_found = 0; }; }

static void put(char *dst, size_t cap, const mpz_class &v) { const std::string s = v.get_str(16); if (s.size() + 1 <= cap) memcpy(dst, s.c_str(), s.size() + 1); else dst[0] = 0; }

extern "C" {

// mix64(seed, idx) and rnd_mpz_bits(2^p - 1, seed0, bits) exactly as the reference defines them (:859-886)
uint64_t ref_ecm_mix64(uint64_t seed, uint64_t idx)
{
	mpz_class N = 1;
#include "_ref/ecm_seed.inc"
	(void)u64_bits; (void)rnd_mpz_bits;
	return mix64(seed, idx);
}

void ref_ecm_rnd_mpz_bits(uint32_t p, uint64_t seed0, unsigned bits, char *out_hex, size_t cap)
{
	mpz_class N = (mpz_class(1) << p) - mpz_class(1);
#include "_ref/ecm_seed.inc"
	(void)u64_bits; (void)mix64;
	put(out_hex, cap, rnd_mpz_bits(N, seed0, bits));
}

// The default (no torsion) curve of one ECM curve index: status 0 = built, 1 = a factor fell out of an inversion
// (factor_hex), -1 = no curve after 256 sigma tries.  Outputs as lower-case hex strings.
int ref_ecm_curve_from_seed(uint32_t p, uint64_t curve_seed, char *aE_hex, char *dE_hex, char *X0_hex, char *Y0_hex, char *sigma_hex,
	char *factor_hex, size_t cap)
{
	mpz_class N = (mpz_class(1) << p) - mpz_class(1);
#include "_ref/ecm_seed.inc"
	(void)u64_bits;
	// the names the cut-out block refers to
	mpz_class result_factor = 0; std::string result_status = "NF", torsion_used;
	uint64_t curves_tested_for_found = 0; const uint64_t c = 0;
	options_stub options; const bool forceSigma = false, te_use_torsion16 = false, te_use_family_iv_163 = false;
	bool built = false, have_sigma_resume = false; mpz_class sigma_resume, aE, dE, X0, Y0;
	engine_stub *eng = nullptr;
	auto write_result = [&]() {}; auto publish_json = [&]() {};
	int status = -1;
	{
#include "_ref/ecm_helpers.inc"
		(void)addm;
		auto body = [&]() -> int {
#include "_ref/ecm_sigma.inc"
			return 2;
		};
		const int r = body();
		(void)r; (void)curves_tested_for_found; (void)have_sigma_resume;
		if (result_status == "found") status = 1; else if (built) status = 0;
	}
	put(aE_hex, cap, aE); put(dE_hex, cap, dE); put(X0_hex, cap, X0); put(Y0_hex, cap, Y0);
	put(sigma_hex, cap, sigma_resume); put(factor_hex, cap, result_factor);
	return status;
}

}
