// The C++ adapter compiled against the REFERENCE'S OWN include/marin/engine.h (-DMARIN_B200_USE_REFERENCE_ENGINE_H
// -I<PrMers>/include): engine_b200 is then a subclass of the reference's `engine`, and the reference's own
// get_mpz / set_mpz / is_equal (default, through get_mpz) / pow / digit code runs on top of the B200 engine.
// Sequence of the reference's tests/test_aevum_reg_adapter.cpp:27-87 (exponent 31, set_mpz included), then bigger
// exponents through the multi-pass kernels with GMP as the arithmetic check.  Built in the container that holds
// /root/reference (the header is not copied anywhere); the binary travels to the GPU box.  GMP: libgmp runtime +
// tests/cpp/gmp_shim/gmp.h (declarations only).
#include <cstdio>
#include <memory>
#include "marin_b200_engine.hpp"

static uint64_t value(engine &eng, engine::Reg reg) { engine::digit d(&eng, reg); uint64_t r = d.res64(); return d.equal_to_Mp() ? 0 : r; }
static int fails = 0;
static void expect(uint64_t got, uint64_t want, const char *label)
{
	if (got != want) { std::printf("FAIL %s: got %llu, expected %llu\n", label, (unsigned long long)got, (unsigned long long)want); ++fails; }
}

int main()
{
	const uint32_t exponent = 31;
	const uint64_t modulus = (uint64_t(1) << exponent) - 1;
	std::unique_ptr<engine> eng(create_gpu_b200(exponent, 8, 0, false));
	eng->set(0, 5); eng->set(1, 7);
	eng->set_multiplicand(2, 1); eng->mul(0, 2);
	expect(value(*eng, 0), 35, "mul");
	eng->square_mul(0, 3);
	expect(value(*eng, 0), 3675, "square_mul");
	eng->add(0, 1); eng->sub_reg(0, 1); eng->sub(0, 2);
	expect(value(*eng, 0), 3673, "add/sub");
	mpz_t z; mpz_init(z); mpz_set_ui(z, modulus + 123);
	eng->set_mpz(3, z);                                        // reduced by the adapter, split by the reference's set_mpz (engine.h:205-232)
	expect(value(*eng, 3), 123, "set_mpz");
	std::vector<char> one(eng->get_register_data_size());
	if (!eng->get_data(one, 3)) { std::puts("FAIL get_data"); ++fails; }
	eng->set(4, uint32_t(0));
	if (!eng->set_data(4, one)) { std::puts("FAIL set_data"); ++fails; }
	expect(value(*eng, 4), 123, "register data");
	if (!eng->is_equal(3, 4) || !eng->engine::is_equal(3, 4)) { std::puts("FAIL is_equal"); ++fails; }   // device compare and the reference's GMP compare
	eng->set(4, 124);
	if (eng->is_equal(3, 4) || eng->engine::is_equal(3, 4)) { std::puts("FAIL !is_equal"); ++fails; }
	eng->set(4, 123);
	std::vector<char> ck(eng->get_checkpoint_size());
	if (!eng->get_checkpoint(ck)) { std::puts("FAIL get_checkpoint"); ++fails; }
	eng->set(0, 1); eng->set(3, 1);
	if (!eng->set_checkpoint(ck)) { std::puts("FAIL set_checkpoint"); ++fails; }
	expect(value(*eng, 0), 3673, "checkpoint r0");
	expect(value(*eng, 3), 123, "checkpoint r3");
	eng->set(5, 9); eng->pow(6, 5, 13);                        // the reference's pow (engine.h:159-170)
	uint64_t e = 1; for (int i = 0; i < 13; ++i) e = (e * 9) % modulus;
	expect(value(*eng, 6), e, "pow");
	engine::digit digits(eng.get(), 3);
	if (digits.get_size() != eng->get_size() || digits.res64() != 123) { std::puts("FAIL digit"); ++fails; }
	eng->sync();

	// larger exponents (single-CTA, three-pass and the specialised n = 2^23 kernels): x -> x^2 * 3 chains checked with
	// GMP through the reference's get_mpz / set_mpz
	for (const uint32_t p : {9941u, 1362763u, 136279841u})
	{
		std::unique_ptr<engine> g(create_gpu_b200(p, 4, 0, false));
		mpz_t m, x, t; mpz_inits(m, x, t, nullptr);
		mpz_ui_pow_ui(m, 2, p); mpz_sub_ui(m, m, 1);
		mpz_ui_pow_ui(x, 3, 211);                               // a few hundred bits to start from
		g->set_mpz(0, x);
		const int steps = (p > 100000000u) ? 6 : 40;
		for (int i = 0; i < steps; ++i)
		{
			g->square_mul(0, 3);
			mpz_mul(t, x, x); mpz_mod(x, t, m);                  // x <- x^2 mod 2^p - 1
			mpz_set_ui(t, 3); mpz_mul(t, x, t); mpz_mod(x, t, m); // x <- 3 x mod 2^p - 1
		}
		g->get_mpz(t, 0);
		if (mpz_cmp(t, x) != 0) { std::printf("FAIL chain p = %u\n", p); ++fails; }
		g->copy(1, 0); g->sub(1, 1);
		if (g->engine::is_equal(0, 1) || g->is_equal(0, 1)) { std::printf("FAIL is_equal p = %u\n", p); ++fails; }
		g->set(2, 1); g->add(1, 2);
		if (!g->engine::is_equal(0, 1) || !g->is_equal(0, 1)) { std::printf("FAIL is_equal (2) p = %u\n", p); ++fails; }
		mpz_clears(m, x, t, nullptr);
	}
	mpz_clear(z);
	if (fails == 0) std::puts("engine_b200 adapter test against the reference engine.h passed");
	return fails;
}
