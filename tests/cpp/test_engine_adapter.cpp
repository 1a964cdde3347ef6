// Conformance test of the C++ adapter (include/marin_b200_engine.hpp) on a real GPU: the sequence of
// the reference's own adapter test, tests/test_aevum_reg_adapter.cpp:27-87, exponent 31, plus a
// Lucas-Lehmer run of M127 (`prmers 127 -ll`, BASELINE config 1) and error behaviour.
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
	std::vector<char> one(eng->get_register_data_size());
	eng->set(3, 123);
	if (!eng->get_data(one, 3)) { std::puts("FAIL get_data"); ++fails; }
	eng->set(4, uint32_t(0));
	if (!eng->set_data(4, one)) { std::puts("FAIL set_data"); ++fails; }
	expect(value(*eng, 4), 123, "register data");
	if (!eng->is_equal(3, 4)) { std::puts("FAIL is_equal"); ++fails; }
	eng->set(4, 124);
	if (eng->is_equal(3, 4)) { std::puts("FAIL !is_equal"); ++fails; }
	eng->set(4, 123);
	std::vector<char> ck(eng->get_checkpoint_size());
	if (!eng->get_checkpoint(ck)) { std::puts("FAIL get_checkpoint"); ++fails; }
	eng->set(0, 1); eng->set(3, 1);
	if (!eng->set_checkpoint(ck)) { std::puts("FAIL set_checkpoint"); ++fails; }
	expect(value(*eng, 0), 3673, "checkpoint r0");
	expect(value(*eng, 3), 123, "checkpoint r3");
	std::vector<char> bad(3);
	if (eng->get_data(bad, 0) || eng->set_checkpoint(bad)) { std::puts("FAIL size mismatch must return false"); ++fails; }
	eng->set(5, 9); eng->pow(6, 5, 13);
	uint64_t e = 1; for (int i = 0; i < 13; ++i) e = (e * 9) % modulus;
	expect(value(*eng, 6), e, "pow");
	engine::digit digits(eng.get(), 3);
	if (digits.get_size() != eng->get_size() || digits.res64() != 123) { std::puts("FAIL digit"); ++fails; }
	eng->sync();
	bool threw = false;
	try { eng->copy(0, 99); } catch (const std::runtime_error &) { threw = true; }
	if (!threw) { std::puts("FAIL out-of-range register must throw"); ++fails; }

	// Lucas-Lehmer M127: s <- s^2 - 2, 125 times from 4, ends at 0 (mod 2^127 - 1)
	std::unique_ptr<engine> ll(create_gpu_b200(127, 2, 0, false));
	ll->set(0, 4);
	for (int i = 0; i < 125; ++i) { ll->square_mul(0); ll->sub(0, 2); }
	engine::digit r(ll.get(), 0);
	if (!(r.equal_to(0) || r.equal_to_Mp())) { std::puts("FAIL LL M127"); ++fails; }

	if (fails == 0) std::puts("engine_b200 adapter test passed");
	return fails;
}
