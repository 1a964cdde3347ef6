// The reference's OWN run-time plugin adapter, src/aevum/EngineAevum.cpp (compiled where it lies, unmodified), loading
// libmarin_b200.so through $AEVUM_ENGINE_LIB and driving it through the Aevum plugin ABI that library also exports:
// the body of the reference's tests/test_aevum_reg_adapter.cpp:27-87 (set / mul / square_mul / add / sub, set_mpz,
// register data, is_equal, checkpoints, pow, digit), with create_aevum_engine in place of engine::create_gpu (the
// factory lives in src/marin/gpu.cpp, which needs OpenCL headers), then a squaring chain at the headline exponent.
// Test infrastructure: built in the container that holds /root/reference; the binary travels to the GPU box.
#include <cstdint>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include "aevum/EngineAevum.hpp"
#include "marin/engine.h"

namespace {
std::uint64_t value(engine& eng, engine::Reg reg) { engine::digit digits(&eng, reg); return digits.res64(); }
void expect(std::uint64_t actual, std::uint64_t expected, const char* label) {
    if (actual != expected) throw std::runtime_error(std::string(label) + ": got " + std::to_string(actual) + ", expected " + std::to_string(expected));
}
}

int main() {
    try {
        constexpr std::uint32_t exponent = 31;
        constexpr std::uint64_t modulus = (std::uint64_t(1) << exponent) - 1;
        std::unique_ptr<engine> eng(create_aevum_engine(exponent, 8, 0, false, ""));
        if (!eng->is_aevum_backend()) return 2;
        eng->set(0, 5); eng->set(1, 7);
        eng->set_multiplicand(2, 1); eng->mul(0, 2);
        expect(value(*eng, 0), 35, "mul");
        eng->square_mul(0, 3);
        expect(value(*eng, 0), 3675, "square_mul");
        eng->add(0, 1); eng->sub_reg(0, 1); eng->sub(0, 2);
        expect(value(*eng, 0), 3673, "add/sub");
        mpz_t z; mpz_init(z); mpz_set_ui(z, modulus + 123);
        eng->set_mpz(3, z); mpz_clear(z);
        expect(value(*eng, 3), 123, "set_mpz");
        std::vector<char> one(eng->get_register_data_size());
        if (!eng->get_data(one, 3)) return 4;
        eng->set(4, std::uint32_t(0));
        if (!eng->set_data(4, one)) return 5;
        expect(value(*eng, 4), 123, "register data");
        if (!eng->is_equal(3, 4)) return 10;
        eng->set(4, 124);
        if (eng->is_equal(3, 4)) return 11;
        eng->set(4, 123);
        std::vector<char> checkpoint(eng->get_checkpoint_size());
        if (!eng->get_checkpoint(checkpoint)) return 6;
        eng->set(0, 1); eng->set(3, 1);
        if (!eng->set_checkpoint(checkpoint)) return 7;
        expect(value(*eng, 0), 3673, "checkpoint r0");
        expect(value(*eng, 3), 123, "checkpoint r3");
        eng->set(5, 9); eng->pow(6, 5, 13);
        std::uint64_t expected = 1;
        for (int i = 0; i < 13; ++i) expected = (expected * 9) % modulus;
        expect(value(*eng, 6), expected, "pow");
        engine::digit digits(eng.get(), 3);
        if (digits.get_size() != eng->get_size() || digits.res64() != 123) return 8;
        eng->sync();
        eng.reset();

        // headline exponent through the same loader: 3 -> 3^(2^40), res64 = 3^(2^40) mod 2^64 only while it fits; compare
        // with an independent GMP chain instead
        const std::uint32_t p = 136279841;
        std::unique_ptr<engine> big(create_aevum_engine(p, 4, 0, false, ""));
        big->set(0, 3);
        for (int i = 0; i < 30; ++i) big->square_mul(0);
        mpz_t m, x, t; mpz_inits(m, x, t, nullptr);
        mpz_ui_pow_ui(m, 2, p); mpz_sub_ui(m, m, 1);
        mpz_set_ui(x, 3);
        for (int i = 0; i < 30; ++i) { mpz_mul(t, x, x); mpz_mod(x, t, m); }
        big->get_mpz(t, 0);
        if (mpz_cmp(t, x) != 0) throw std::runtime_error("squaring chain at M136279841 differs from GMP");
        mpz_clears(m, x, t, nullptr);
        std::cout << "reference plugin loader (engine_aevum) over libmarin_b200.so: test passed" << std::endl;
        return 0;
    } catch (const std::exception& e) { std::cout << "FAIL " << e.what() << std::endl; return 1; }
}
