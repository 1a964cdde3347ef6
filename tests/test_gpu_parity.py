"""GPU tier: the CUDA engine, called through the C ABI (libmarin_b200.so via prmers_b200.Engine),
against the CPU oracle, big-int arithmetic and the reference's known-answer vectors.
Bit-exact: integer arithmetic, no tolerance anywhere."""
import os
import random

import numpy as np
import pytest

from oracle import marin_oracle as mo
from oracle import mersenne_ref as mr

pytestmark = pytest.mark.gpu


def _engine(p, regs, force_S=0, force_R2=0):
    import prmers_b200 as pb
    return pb.Engine(p, regs, force_S=force_S, force_R2=force_R2)


def val(e, r):
    return e.get_mpz(r)


CASES = [(31, 0), (89, 0), (127, 0), (521, 0), (701, 0), (1279, 0), (3217, 0), (9941, 0), (86243, 0), (216091, 0),
         (9941, 16), (86243, 8), (86243, 64), (86243, 256), (216091, 64), (216091, 128), (1362763, 0), (2500009, 0),
         (5000011, 0), (20000003, 0), (311111, 0), (620003, 0), (1250003, 0), (1560007, 0), (9999991, 0), (12500003, 0)]


@pytest.mark.parametrize("p,force_S", CASES)
def test_register_ops_match_bigint(p, force_S):
    """Same sequence as tests/test_emu_kernels.py, now on the real kernels (engine.h:45-74;
    the conformance shape of tests/test_aevum_reg_adapter.cpp:27-87)."""
    e = _engine(p, 6, force_S)
    assert e.get_size() == mo.transform_size(p)
    m = mr.mp(p)
    rng = random.Random(p * 7 + force_S)
    x, y = rng.randrange(m), rng.randrange(m)
    e.set_mpz(0, x); e.set_mpz(1, y)
    assert val(e, 0) == x and val(e, 1) == y
    e.set_multiplicand(2, 1)
    assert val(e, 1) == y
    a = min(3, e.max_multiplier())
    M = lambda u, v: mr.mulmod(u, v, p)
    e.mul(0, 2, a); x = M(M(x, y), a)
    assert val(e, 0) == x
    e.mul(0, 2); x = M(x, y)
    assert val(e, 0) == x
    e.square_mul(0, a); x = M(M(x, x), a)
    assert val(e, 0) == x
    e.square_mul(0); x = M(x, x)
    assert val(e, 0) == x
    e.add(0, 1); x = (x + y) % m
    assert val(e, 0) == x
    e.sub_reg(0, 1); e.sub_reg(0, 1); x = (x - 2 * y) % m
    assert val(e, 0) == x
    e.sub(0, 2); x = (x - 2) % m
    assert val(e, 0) == x
    e.add(0, 0); x = 2 * x % m
    assert val(e, 0) == x
    e.copy(3, 0); e.set(4, 0); e.sub(4, 5)
    assert val(e, 3) == x and val(e, 4) == (m - 5) % m
    e.set(5, 0xFFFFFFFF)
    assert val(e, 5) == 0xFFFFFFFF % m
    e.set_multiplicand(3, 3)
    e.set(4, 1); e.mul(4, 3)
    assert val(e, 4) == x
    e.square_mul(0); x = M(x, x)
    assert val(e, 0) == x
    e.close()


@pytest.mark.parametrize("p,force_S", [(127, 0), (701, 0), (9941, 0), (9941, 16), (86243, 0), (86243, 64), (216091, 0),
                                       (216091, 64), (1362763, 0), (311111, 0), (620003, 0), (1250003, 0), (1560007, 0),
                                       (9999991, 0), (12500003, 0), (20996011, 0)])
def test_digits_bit_exact_vs_oracle(p, force_S):
    e, o = _engine(p, 2, force_S), mo.OracleEngine(p, 2)
    e.set(0, 3); o.set(0, 3)
    for i in range(60):
        a = 3 if (i % 3 == 0 and e.max_multiplier() >= 3) else 1
        e.square_mul(0, a); o.square_mul(0, a)
    assert np.array_equal(e.get(0), o.get_digits(0))
    e.sub(0, 2); o.sub(0, 2)
    assert np.array_equal(e.get(0), o.get_digits(0))
    e.close()


def test_api_conformance_like_reference_adapter_test():
    """tests/test_aevum_reg_adapter.cpp:27-87, exponent 31."""
    import prmers_b200 as pb
    modulus = (1 << 31) - 1
    e = pb.create_gpu(31, 8, 0, False)
    e.set(0, 5); e.set(1, 7)
    e.set_multiplicand(2, 1); e.mul(0, 2)
    assert val(e, 0) == 35
    e.square_mul(0, 3)
    assert val(e, 0) == 3675
    e.add(0, 1); e.sub_reg(0, 1); e.sub(0, 2)
    assert val(e, 0) == 3673
    e.set_mpz(3, modulus + 123)
    assert val(e, 3) == 123
    one = bytearray(e.get_register_data_size())
    assert e.get_data(one, 3)
    e.set(4, 0)
    assert e.set_data(4, one)
    assert val(e, 4) == 123
    assert e.is_equal(3, 4)
    e.set(4, 124)
    assert not e.is_equal(3, 4)
    e.set(4, 123)
    ck = bytearray(e.get_checkpoint_size())
    assert e.get_checkpoint(ck)
    e.set(0, 1); e.set(3, 1)
    assert e.set_checkpoint(ck)
    assert val(e, 0) == 3673 and val(e, 3) == 123
    assert not e.get_data(bytearray(3), 0) and not e.set_checkpoint(bytearray(5))
    e.set(5, 9); e.pow(6, 5, 13)
    assert val(e, 6) == pow(9, 13, modulus)
    d = e.digit(3)
    assert d.get_size() == e.get_size() and d.res64() == 123 and d.equal_to(123) and not d.equal_to_Mp()
    e.sync()
    with pytest.raises(pb.MarinError):
        e.copy(0, 99)
    with pytest.raises(pb.MarinError):
        e.square_mul(0, 0)
    e.close()


def test_fused_helpers_match_default_compositions():
    """engine.h:65-131 defaults define the semantics of mul_add / addsub / *_copy."""
    p = 9941
    e = _engine(p, 10)
    m = mr.mp(p)
    rng = random.Random(5)
    a, b, c = (rng.randrange(m) for _ in range(3))
    e.set_mpz(0, a); e.set_mpz(1, b); e.set_mpz(2, c)
    e.addsub(3, 4, 0, 1)
    assert val(e, 3) == (a + b) % m and val(e, 4) == (a - b) % m
    e.addsub_copy(3, 4, 5, 6, 0, 1)
    assert val(e, 5) == (a + b) % m and val(e, 6) == (a - b) % m
    e.set_multiplicand(7, 1)
    e.mul_add(0, 7, 2, 3)
    assert val(e, 0) == (a * b * 3 + c) % m
    e.square_mul_copy(2, 8, 3)
    assert val(e, 2) == val(e, 8) == c * c * 3 % m
    e.mul_copy(2, 7, 9)
    assert val(e, 2) == val(e, 9) == c * c * 3 * b % m
    e.close()


@pytest.mark.parametrize("p,force_S,force_R2", [(127, 0, 0), (701, 0, 0), (9941, 0, 0), (86243, 0, 0), (86243, 64, 0), (216091, 64, 0),
                                                (86243, 8, 16), (1362763, 0, 0)])
def test_fused_kernels_and_device_canonical_export(p, force_S, force_R2):
    """addsub(+copy), square_mul_copy, mul_copy as single passes; canonical carry and is_equal on the device."""
    from tests.test_emu_kernels import fused_ops_script
    e = _engine(p, 10, force_S, force_R2)
    assert fused_ops_script(e, p, lambda eng, r, x, wd: eng.set_mpz(r, x), val, lambda eng, l, r: eng.is_equal(l, r))
    e.close()


def test_fused_kernels_full_size():
    """The same at the headline size (n = 2^23, specialised kernels) and at n = 5 * 2^21 (config 4): the fused kernels
    against the reference's default compositions (engine.h:65-131), canonical digits compared on the device and digit
    for digit on the host -- no big-integer conversions at this size."""
    for p in (136279841, 205271257):
        e = _engine(p, 12)
        e.set(0, 3); e.set(1, 5)
        for _ in range(30):
            e.square_mul(0); e.square_mul(1, 3)
        # compositions into 2..5
        e.copy(2, 0); e.copy(3, 0); e.add(2, 1); e.sub_reg(3, 1)
        e.addsub(4, 5, 0, 1)
        assert e.is_equal(2, 4) and e.is_equal(3, 5) and not e.is_equal(4, 5)
        assert np.array_equal(e.get(2), e.get(4)) and np.array_equal(e.get(3), e.get(5))
        e.copy(6, 0); e.copy(7, 1)
        e.addsub_copy(6, 7, 8, 9, 6, 7)                       # in place, with copies
        assert e.is_equal(6, 2) and e.is_equal(7, 3) and e.is_equal(8, 2) and e.is_equal(9, 3)
        e.copy(2, 0); e.square_mul(2, 3); e.copy(3, 2)
        e.copy(4, 0)
        for _ in range(3):
            e.square_mul(10)                                  # another register parked in between
        e.square_mul_copy(4, 5, 3)
        assert e.is_equal(4, 2) and e.is_equal(5, 2) and np.array_equal(e.get(5), e.get(3))
        e.set_multiplicand(11, 1)
        e.copy(2, 0); e.mul(2, 11, 3); e.copy(3, 2)
        e.copy(4, 0); e.mul_copy(4, 11, 5, 3)
        assert e.is_equal(4, 2) and e.is_equal(5, 3)
        e.copy(2, 0); e.mul(2, 11); e.add(2, 1)
        e.copy(4, 0); e.mul_add(4, 11, 1)
        assert e.is_equal(4, 2)
        e.close()


@pytest.mark.parametrize("p", [136279841, 205271257, 100000007, 77232917, 50000017, 37156667, 25000009, 20996011,
                               12500003, 9999991, 1250003, 620003, 311111])
def test_specialised_kernels_match_generic_kernels_bit_for_bit(p):
    """marin_fast.cuh against the generic phase functions: identical raw register images (MARIN_B200_NO_FAST=1), at every
    transform size whose planned split uses a specialised kernel (all six row shapes of the contiguous pass; the R = 1024
    strided kernels from n = 2^21 / 5 * 2^19 up)."""
    import prmers_b200 as pb
    imgs = []
    for fast in (True, False):
        if not fast:
            os.environ["MARIN_B200_NO_FAST"] = "1"
        try:
            e = pb.Engine(p, 4)
        finally:
            os.environ.pop("MARIN_B200_NO_FAST", None)
        out = []

        def raw(r):
            b = bytearray(e.get_register_data_size())
            assert e.get_data(b, r)
            return bytes(b)
        e.set(0, 3)
        for _ in range(30):
            e.square_mul(0)
        e.copy(1, 0)
        for _ in range(5):
            e.square_mul(0)
        out.append(raw(0))
        e.square_mul(0, 3); out.append(raw(0))
        e.set_multiplicand(2, 1); out.append(raw(2))
        e.mul(0, 2, 1); out.append(raw(0))
        e.mul(1, 2, 3); out.append(raw(1))
        e.square_mul_copy(1, 3, 1); out.append(raw(1)); out.append(raw(3))
        imgs.append(out)
        e.close()
    for i, (x, y) in enumerate(zip(*imgs)):
        assert x == y, f"step {i}"


def test_m127_lucas_lehmer_and_m86243_prp():
    """BASELINE config 1: LL M127 and PRP p = 86243, bit-exact residue plumbing."""
    e = _engine(127, 2)
    e.set(0, 4)
    for _ in range(125):
        e.square_mul(0); e.sub(0, 2)
    assert val(e, 0) == 0
    e.close()
    p = 86243
    e = _engine(p, 8)
    assert e.get_size() == 4096
    e.set(0, 3)
    e.square_mul_repeat(0, p)
    d = e.digit(0)
    assert d.equal_to(9)                                       # prime: RunPrpOrLlMarin.cpp:453
    _, r = True, val(e, 0) * pow(9, -1, mr.mp(p)) % mr.mp(p)
    assert mr.res64_hex(r) == "0000000000000001"
    e.close()


M11213_RES64 = {1000: "FBA631FBCB73A011", 2000: "F01283650C4A1491", 3000: "7E79193B757010B7",
                4000: "31482E4D80FE99BB", 5000: "973B76BACF73BBEF", 6000: "8CFFB332495FC320",
                7000: "98080C76DF068843", 8000: "8FDA516F885D3FEE", 9000: "2AADBC4F1E318E92",
                10000: "0A4AAF339C8B290C", 11000: "A1F26F470CFE412D"}


@pytest.mark.parametrize("force_S", [0, 16])
def test_m11213_interim_res64_golden(force_S):
    """unit_tests.sh:151-187 (values are 3^(2^iter) mod M11213, low 64 bits)."""
    e = _engine(11213, 2, force_S)
    e.set(0, 3)
    for it in range(1, 11214):
        e.square_mul(0)
        if it in M11213_RES64:
            assert f"{e.digit(0).res64():016X}" == M11213_RES64[it], it
    assert e.digit(0).equal_to(9)
    e.close()


def test_m100003_res64_res2048_golden():
    """unit_tests.sh:137-148: res64 1CF45E9503C71FD6 and the full res2048."""
    from tests.test_oracle_golden import M100003_RES2048
    p = 100003
    e = _engine(p, 2)
    e.set(0, 3)
    e.square_mul_repeat(0, p)
    m = mr.mp(p)
    r = val(e, 0) * pow(9, -1, m) % m
    assert mr.res64_hex(r) == "1CF45E9503C71FD6"
    assert mr.res2048_hex(r) == M100003_RES2048
    e.close()


def test_pm1_stage1_factors_golden():
    """unit_tests.sh:54-69,204-213: fast3 stage 1 (RunPM1.cpp:7075-7135) finds the known factors."""
    from math import gcd
    for p, b1, factor in [(269, 2141, 13822297), (541, 899, 4312790327), (367, 11981, 646300400639),
                          (2151, 256, 327405968242246366421788399), (544139, 3, 22853839)]:
        E = mr.pm1_exponent(p, b1)
        e = _engine(p, 2)
        e.set(0, 1)
        for bit in bin(E)[2:]:
            e.square_mul(0, 3 if bit == "1" else 1)
        x = val(e, 0)
        assert gcd(x - 1, mr.mp(p)) == factor
        e.close()


def test_extreme_digits_do_not_overflow():
    for p, fs in [(127, 0), (86243, 0), (86243, 64), (216091, 64), (1362763, 0), (104857601, 0)]:
        e = _engine(p, 3, fs)
        n = e.get_size()
        wd = mo.widths(n, p)
        ones = (np.uint64(1) << wd.astype(np.uint64)) - np.uint64(1)
        e.set_digits(0, ones | (wd.astype(np.uint64) << np.uint64(32)))
        e.square_mul(0)
        assert val(e, 0) == 0
        e.set(1, 0); e.sub(1, 1)                               # 2^p - 2 = -1
        e.square_mul(1)
        assert val(e, 1) == 1
        e.set_multiplicand(2, 1); e.mul(1, 2)
        assert val(e, 1) == 1
        e.close()


# ---- BASELINE.json full size: n = 2^23 (M136279841) and n = 5 * 2^21 (p = 205271257) ----

@pytest.mark.parametrize("p,n", [(136279841, 1 << 23), (205271257, 5 << 21), (100000007, 5 << 20), (77232917, 1 << 22),
                                 (50000017, 5 << 19), (37156667, 1 << 21), (25000009, 5 << 18)])
def test_full_size_squarings_match_oracle_and_identities(p, n):
    e = _engine(p, 6)
    assert e.get_size() == n
    o = mo.OracleEngine(p, 1)
    e.set(0, 3); o.set(0, 3)
    for _ in range(3):                                         # 3^(2^3) is tiny; digits are exact
        e.square_mul(0); o.square_mul(0)
    assert np.array_equal(e.get(0), o.get_digits(0))
    # fill with a pseudo-random canonical residue and compare one more squaring digit for digit
    rng = np.random.default_rng(p)
    wd = mo.widths(n, p).astype(np.uint64)
    x = (rng.integers(0, 1 << 32, size=n, dtype=np.uint64) & ((np.uint64(1) << wd) - np.uint64(1))) | (wd << np.uint64(32))
    e.set_digits(0, x); o.set_digits(0, x)
    assert np.array_equal(e.get(0), o.get_digits(0))
    e.copy(1, 0); e.copy(2, 0)
    e.square_mul(0, 3); o.square_mul(0, 3)
    got = e.get(0)
    assert np.array_equal(got, o.get_digits(0))
    # size-independent identities on the device: x*x via mul equals square_mul; (x+x)^2 = 4 x^2
    e.set_multiplicand(3, 1); e.mul(2, 3, 3)
    assert e.is_equal(2, 0)
    e.add(1, 1); e.square_mul(1, 3)                            # 3 (2x)^2 = 12 x^2
    e.copy(4, 0); e.add(4, 4); e.add(4, 4)                     # 4 * (3 x^2)
    assert e.is_equal(1, 4)
    e.sub_reg(4, 1)
    assert val(e, 4) == 0
    e.close()


def test_interim_res64_golden_full_size():
    """tests/golden/prp_interim_res64.json (GMP big integers, tools/make_golden_prp.py): the type-1 PRP state after i
    squarings of 3 at the BASELINE.json sizes, read back in the middle of the chain (the digits are parked in
    transform form between squarings and materialised for each read)."""
    import json
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "prp_interim_res64.json")))["res64"]
    for p, marks in gold.items():
        e = _engine(int(p), 2)
        e.set(0, 3)
        done = 0
        for it in sorted(int(k) for k in marks):
            e.square_mul_repeat(0, it - done); done = it
            assert e.digit(0).res64() == int(marks[str(it)], 16), (p, it)
        e.close()


@pytest.mark.parametrize("p,force_S,force_R2", [(86243, 64, 0), (216091, 64, 0), (86243, 8, 16), (216091, 128, 0)])
def test_chain_state_is_materialised_on_demand(p, force_S, force_R2):
    """A chain, then a product on another register, then the chain again (the parked transform must not survive the
    product: regression), reads, copies, linear ops and overwrites of a parked register."""
    from tests.test_emu_kernels import chain_state_script
    e = _engine(p, 6, force_S, force_R2)
    assert not e.is_single_cta()
    assert chain_state_script(e, p, lambda eng, r, x, wd: eng.set_mpz(r, x), val)
    e.close()


def test_long_chain_gerbicz_identity_full_size():
    """Gerbicz-Li style check (RunPrpOrLlMarin.cpp:338-369) at n = 2^23: with d_k the product of the
    residues at block ends, d_{k+1} == d_k^(2^B) * 3 ... here in its simplest size-independent
    form: (x^(2^B)) computed in one chain equals the chain run in two halves from a checkpoint."""
    p = 136279841
    e = _engine(p, 4)
    e.set(0, 3)
    e.square_mul_repeat(0, 200)
    e.copy(1, 0)
    ck = bytearray(e.get_register_data_size())
    assert e.get_data(ck, 1)
    e.square_mul_repeat(0, 300)
    e.set(2, 0)
    assert e.set_data(2, ck)
    e.square_mul_repeat(2, 300)
    assert e.is_equal(0, 2)
    # multiplicative check: (a*b)^2 == a^2 * b^2
    e.copy(3, 0); e.set_multiplicand(1, 1); e.mul(3, 1)         # r3 = a*b
    e.square_mul(3)
    e.square_mul(0)                                            # a^2
    e.set(2, 0); e.set_data(2, ck); e.square_mul(2)            # b^2
    e.set_multiplicand(2, 2); e.mul(0, 2)
    assert e.is_equal(0, 3)
    e.close()


@pytest.mark.parametrize("p,force_S,force_R2", [(9941, 4, 2), (86243, 8, 16), (216091, 16, 8), (1362763, 64, 8), (20000003, 4096, 4)])
def test_two_level_strided_passes_forced(p, force_S, force_R2):
    e = _engine(p, 4, force_S, force_R2)
    m = mr.mp(p)
    rng = random.Random(p + force_R2)
    x, y = rng.randrange(m), rng.randrange(m)
    M = lambda u, v: mr.mulmod(u, v, p)
    e.set_mpz(0, x); e.set_mpz(1, y)
    for a in (1, 3, 1, 1):
        e.square_mul(0, a); x = M(M(x, x), a)
        assert val(e, 0) == x
    e.set_multiplicand(2, 1); e.mul(0, 2); x = M(x, y)
    assert val(e, 0) == x
    e.sub(0, 2); e.square_mul(0); x = M((x - 2) % m, (x - 2) % m)
    assert val(e, 0) == x
    e.close()


@pytest.mark.parametrize("p", [300000007, 400000009, 700000001, 1100000011, 2200000007])
def test_large_transforms_two_level(p):
    """n = 2^24 (p = 3.0e8: rows of 8192 pairs, three passes), 2^25 (4.0e8), 5 * 2^23 (7.0e8), 2^26 (1.1e9) (1024 x 4 rows:
    the specialised first level and the streaming second level) and 5 * 2^25 (2.2e9: the generic five-pass kernels).  Exact small powers (3^(2^k) fits a few digits), then size-independent identities on a
    dense pseudo-random residue, and one squaring against GMP."""
    e = _engine(p, 5)
    n = e.get_size()
    assert n == mo.transform_size(p) and n > (1 << 23)
    e.set(0, 3)
    for k in range(1, 7):
        e.square_mul(0)
        assert val(e, 0) == 3 ** (2 ** k)
    rng = np.random.default_rng(p)
    wd = mo.widths(n, p).astype(np.uint64)
    x = (rng.integers(0, 1 << 32, size=n, dtype=np.uint64) & ((np.uint64(1) << wd) - np.uint64(1))) | (wd << np.uint64(32))
    e.set_digits(0, x)
    assert np.array_equal(e.get(0), x)
    xv = val(e, 0)
    e.copy(1, 0); e.copy(2, 0)
    e.square_mul(0, 3)
    if p < 500000000:
        assert val(e, 0) == mr.mulmod(mr.mulmod(xv, xv, p), 3, p)
    e.set_multiplicand(3, 1); e.mul(2, 3, 3)
    assert e.is_equal(2, 0)                                      # x*x via mul == square_mul
    e.add(1, 1); e.square_mul(1, 3)                              # 3 (2x)^2 = 4 * (3 x^2)
    e.copy(4, 0); e.add(4, 4); e.add(4, 4)
    assert e.is_equal(1, 4)
    e.square_mul(0); e.square_mul(0); e.square_mul(4); e.square_mul(4)   # chained path: (4 y)^4 = 256 y^4
    e.set(3, 256); e.set_multiplicand(3, 3); e.mul(0, 3)
    assert e.is_equal(0, 4)
    e.close()


def test_cpp_adapter_conformance():
    """include/marin_b200_engine.hpp: the C++ `engine` subclass over the C ABI, built with g++ and run
    on the GPU (sequence of the reference's tests/test_aevum_reg_adapter.cpp)."""
    import subprocess
    import prmers_b200 as pb
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pb.load_library()
    exe = os.path.join(root, "tests", "cpp", "test_engine_adapter")
    libdir = os.path.join(root, "prmers_b200")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-std=c++17", "-O1", "-I", os.path.join(root, "include"), "-o", exe,
                    os.path.join(root, "tests", "cpp", "test_engine_adapter.cpp"),
                    "-L", libdir, "-l:libmarin_b200.so", f"-Wl,-rpath,{libdir}"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "adapter test passed" in r.stdout, r.stdout + r.stderr


def test_cpp_adapter_against_reference_engine_header():
    """engine_b200 as a subclass of the reference's OWN `engine` (include/marin/engine.h, compiled in the container that
    holds /root/reference by __graft_entry__.build(); only the binary travels): the reference's get_mpz / set_mpz /
    default is_equal / pow / digit run over the B200 engine, exponents 31 ... 136279841, GMP as the arithmetic check."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "cpp", "test_engine_adapter_refhdr")
    if not os.path.exists(exe):
        pytest.skip("tests/cpp/test_engine_adapter_refhdr not built (needs /root/reference: python __graft_entry__.py)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "against the reference engine.h passed" in r.stdout, r.stdout + r.stderr


def test_reference_plugin_loader_drives_the_engine():
    """INTEGRATION.md route B: the reference's OWN run-time adapter (src/aevum/EngineAevum.cpp: dlopen of $AEVUM_ENGINE_LIB,
    `engine_aevum`), compiled unmodified where it lies, loads libmarin_b200.so through the Aevum plugin ABI the library
    also exports and passes the reference's adapter test sequence (tests/test_aevum_reg_adapter.cpp:27-87) plus a
    squaring chain at M136279841 against GMP."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "cpp", "test_reference_plugin_loader")
    if not os.path.exists(exe):
        pytest.skip("tests/cpp/test_reference_plugin_loader not built (needs /root/reference: python __graft_entry__.py)")
    env = dict(os.environ, AEVUM_ENGINE_LIB=os.path.join(root, "prmers_b200", "libmarin_b200.so"))
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and "test passed" in r.stdout, r.stdout + r.stderr
