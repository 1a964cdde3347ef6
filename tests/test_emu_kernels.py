"""CPU tier: the product's kernel phase functions and host logic (prmers_b200/csrc) executed by the
thread-by-thread emulator in tests/emu, checked against the oracle and big-int arithmetic.  This
validates indexing / twiddles / carries without a GPU; the GPU tier (test_gpu_parity.py) repeats the
same comparisons through the C-ABI on the real kernels."""
import random

import numpy as np
import pytest

from oracle import marin_oracle as mo
from oracle import mersenne_ref as mr
from tests.emu.emu import EmuEngine, lib


def val(e, r):
    v = mo.digits_to_int(e.get_digits(r))
    return v % ((1 << e.p) - 1)


def put(e, r, x, wd):
    e.set_digits(r, mo.int_to_digits(x, wd))


# (exponent, forced contiguous block S; 0 = planner's choice)
CASES = [(31, 0), (89, 0), (127, 0), (521, 0), (701, 0), (1279, 0), (3217, 0), (9941, 0), (86243, 0), (216091, 0),
         (9941, 16), (86243, 8), (86243, 64), (86243, 256), (216091, 64), (216091, 128), (1362763, 0)]


@pytest.mark.parametrize("p,force_S", CASES)
def test_register_ops_match_bigint(p, force_S):
    e = EmuEngine(p, 6, force_S)
    assert e.n == mo.transform_size(p)
    wd = mo.widths(e.n, p)
    m = mr.mp(p)
    rng = random.Random(p * 7 + force_S)
    x, y = rng.randrange(m), rng.randrange(m)
    put(e, 0, x, wd); put(e, 1, y, wd)
    assert val(e, 0) == x and val(e, 1) == y
    e.set_multiplicand(2, 1)
    assert val(e, 1) == y                                   # src untouched
    a = min(3, e.a_max)
    e.mul(0, 2, a); x = x * y * a % m
    assert val(e, 0) == x
    e.mul(0, 2); x = x * y % m                                # multiplicand reusable
    assert val(e, 0) == x
    e.square_mul(0, a); x = x * x * a % m
    assert val(e, 0) == x
    e.square_mul(0); x = x * x % m
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
    e.set_multiplicand(3, 3)                                  # in place (engine.h:162)
    e.set(4, 1); e.mul(4, 3)
    assert val(e, 4) == x
    e.square_mul(0); x = x * x % m                            # still fine after the linear ops
    assert val(e, 0) == x


def chain_state_script(e, p, put, val):
    """Squaring chains leave the register's digits parked in transform form (no write-back); every other way of
    looking at or changing the register must see the finished digits.  Shared by the emulated and the GPU tier."""
    m = (1 << p) - 1
    wd = mo.widths(e.get_size(), p)
    rng = random.Random(p)
    x, y, z = rng.randrange(m), rng.randrange(m), rng.randrange(m)
    put(e, 0, x, wd); put(e, 1, y, wd); put(e, 4, z, wd)
    for _ in range(4):
        e.square_mul(0); x = x * x % m
    assert val(e, 0) == x                                     # get after a chain
    for _ in range(3):
        e.square_mul(0); x = x * x % m                        # the chain goes on after the read
    e.copy(2, 0)                                              # copy out of a parked register
    assert val(e, 2) == x
    for _ in range(3):
        e.square_mul(0); x = x * x % m
    for _ in range(3):
        e.square_mul(1); y = y * y % m                        # another register squared while 0 is parked
    assert val(e, 0) == x and val(e, 1) == y
    for _ in range(3):
        e.square_mul(0); x = x * x % m
    e.set_multiplicand(3, 4); e.mul(4, 3); z = z * z % m      # mul borrows the parking buffer
    assert val(e, 4) == z
    for _ in range(3):
        e.square_mul(0); x = x * x % m
    e.add(0, 1); x = (x + y) % m                              # linear op on a parked register
    for _ in range(3):
        e.square_mul(0); x = x * x % m
    e.set_multiplicand(3, 0)                                  # parked register as the source of a multiplicand
    e.square_mul(0); x2 = x * x % m
    e.mul(1, 3); y = y * x % m
    assert val(e, 0) == x2 and val(e, 1) == y
    for _ in range(3):
        e.square_mul(0); x2 = x2 * x2 % m
    e.set(0, 7)                                               # overwritten while parked
    e.square_mul(0)
    assert val(e, 0) == 49 % m
    return True


def fused_ops_script(e, p, put, val, is_equal=None):
    """The one-pass kernels behind addsub / addsub_copy / square_mul_copy / mul_copy (marin.cl:1731-2154) against the
    reference's default compositions (engine.h:65-131), including the in-place calls the Edwards schedules make, a
    squaring chain parked before them, and the canonical export / comparison corner cases (carry that ripples through
    runs of all-ones digits, 2^p - 1 == 0).  Shared by the emulated and the GPU tier; needs 10 registers."""
    m = (1 << p) - 1
    wd = mo.widths(e.get_size(), p)
    rng = random.Random(p + 1)
    a, b, c = (rng.randrange(m) for _ in range(3))
    put(e, 0, a, wd); put(e, 1, b, wd); put(e, 2, c, wd)
    e.addsub(3, 4, 0, 1)
    assert val(e, 3) == (a + b) % m and val(e, 4) == (a - b) % m and val(e, 0) == a and val(e, 1) == b
    e.addsub(0, 1, 0, 1)                                      # in place: (a, b) <- (a + b, a - b)
    a, b = (a + b) % m, (a - b) % m
    assert val(e, 0) == a and val(e, 1) == b
    e.addsub(1, 0, 0, 1)                                      # crossed outputs
    a, b = (a - b) % m, (a + b) % m
    assert val(e, 0) == a and val(e, 1) == b
    e.addsub_copy(3, 4, 5, 6, 0, 1)
    assert val(e, 3) == val(e, 5) == (a + b) % m and val(e, 4) == val(e, 6) == (a - b) % m
    e.addsub_copy(0, 1, 5, 6, 0, 1)                           # in place with copies
    a, b = (a + b) % m, (a - b) % m
    assert val(e, 0) == val(e, 5) == a and val(e, 1) == val(e, 6) == b
    amul = min(3, e.a_max) if hasattr(e, "a_max") else 3
    for _ in range(3):
        e.square_mul(2); c = c * c % m                        # leaves register 2 parked on multi-pass engines
    e.square_mul_copy(2, 8, amul); c = c * c * amul % m
    assert val(e, 2) == val(e, 8) == c
    e.square_mul(2); c = c * c % m                            # and the chain restarts cleanly
    assert val(e, 2) == c and val(e, 8) != c
    e.set_multiplicand(7, 1)
    e.mul_copy(2, 7, 9, amul); c = c * b * amul % m
    assert val(e, 2) == val(e, 9) == c and val(e, 1) == b
    e.mul_add(2, 7, 0); c = (c * b + a) % m                   # product and sum carried together
    assert val(e, 2) == c and val(e, 0) == a
    e.mul_add(2, 7, 0, amul); c = (c * b * amul + a) % m
    assert val(e, 2) == c
    # canonical export: all ones is zero; all ones plus one ripples a carry through every run and wraps (2^p = 1)
    put(e, 3, 0, wd)
    ones = mo.int_to_digits(0, wd)
    for k in range(len(ones)):
        ones[k] = (int(ones[k]) & ~0xffffffff) | ((1 << int(wd[k])) - 1)
    e.set_digits(4, ones)
    assert val(e, 4) == 0
    if is_equal:
        assert is_equal(e, 3, 4) and is_equal(e, 4, 3) and is_equal(e, 4, 4) and not is_equal(e, 3, 0)
        e.copy(5, 0)
        assert is_equal(e, 5, 0)
        e.add(5, 3)                                           # + 0: another semi-carried image of the same value
        assert is_equal(e, 5, 0)
    e.set(5, 1)
    e.add(4, 5)                                               # (2^p - 1) + 1 = 2^p = 1
    assert val(e, 4) == 1
    e.set_digits(4, ones); e.set(5, 2**32 - 1 if p > 40 else 5)
    e.add(4, 5)
    assert val(e, 4) == (2**32 - 1 if p > 40 else 5) % m
    return True


@pytest.mark.parametrize("p,force_S", [(127, 0), (701, 0), (9941, 0), (86243, 0), (86243, 64), (216091, 64)])
def test_fused_ops_and_canonical_corner_cases(p, force_S):
    e = EmuEngine(p, 10, force_S)
    assert fused_ops_script(e, p, put, val, lambda eng, l, r: eng.is_equal(l, r))


@pytest.mark.parametrize("p,force_S,force_R2", [(86243, 64, 0), (216091, 64, 0), (86243, 8, 16), (216091, 128, 0)])
def test_chain_state_is_materialised_on_demand(p, force_S, force_R2):
    e = EmuEngine(p, 6, force_S, force_R2)
    assert not e.small
    assert chain_state_script(e, p, put, val)


@pytest.mark.parametrize("p,force_S", [(127, 0), (9941, 0), (9941, 16), (86243, 64), (216091, 64)])
def test_digits_bit_exact_vs_oracle(p, force_S):
    """Canonical digits (value | width << 32) equal the oracle's digit for digit."""
    e, o = EmuEngine(p, 2, force_S), mo.OracleEngine(p, 2)
    e.set(0, 3); o.set(0, 3)
    for i in range(40):
        a = 3 if (i % 3 == 0 and e.a_max >= 3) else 1
        e.square_mul(0, a); o.square_mul(0, a)
    assert np.array_equal(e.get_digits(0), o.get_digits(0))
    e.sub(0, 2); o.sub(0, 2)
    assert np.array_equal(e.get_digits(0), o.get_digits(0))


def test_extreme_digits_do_not_overflow():
    """All-ones digits (value 2^p - 1 == 0) and the largest canonical values: the headroom rule
    n (2^(w+1) - 1)^2 < P of ibdwt.h:28-30 must hold through the transform and carry."""
    for p, fs in [(127, 0), (86243, 0), (86243, 64), (216091, 64), (161051, 0)]:
        e = EmuEngine(p, 3, fs)
        wd = mo.widths(e.n, p)
        m = mr.mp(p)
        ones = (np.uint64(1) << wd.astype(np.uint64)) - np.uint64(1)
        e.set_digits(0, ones | (wd.astype(np.uint64) << np.uint64(32)))
        e.square_mul(0)
        assert val(e, 0) == 0
        x = m - 1
        put(e, 1, x, wd)
        e.square_mul(1)
        assert val(e, 1) == 1
        e.set_multiplicand(2, 1); e.mul(1, 2)
        assert val(e, 1) == 1


def test_m127_lucas_lehmer_and_prp():
    e = EmuEngine(127, 2)
    e.set(0, 4)
    for _ in range(125):
        e.square_mul(0); e.sub(0, 2)
    assert val(e, 0) == 0
    e.set(1, 3)
    for _ in range(127):
        e.square_mul(1)
    assert val(e, 1) == 9


def test_m11213_interim_res64_first_thousand():
    e = EmuEngine(11213, 1)
    e.set(0, 3)
    for _ in range(1000):
        e.square_mul(0)
    assert mr.res64_hex(val(e, 0)) == "FBA631FBCB73A011"        # unit_tests.sh:164-187


def test_planner_and_limits():
    L = lib()
    for p in (31, 127, 701, 86243, 1362763, 136279841, 175000001, 205271257):
        assert L.emu_transform_size(p) == mo.transform_size(p)
    e = EmuEngine(127, 2)
    with pytest.raises(RuntimeError):
        e.square_mul(0, 0)
    with pytest.raises(RuntimeError):
        e.copy(0, 9)
    with pytest.raises(ValueError):
        EmuEngine(2, 1)


@pytest.mark.parametrize("p,force_S,force_R2", [(9941, 4, 2), (86243, 8, 16), (86243, 16, 4), (216091, 16, 8), (1362763, 64, 8)])
def test_two_level_strided_passes(p, force_S, force_R2):
    """Transforms with more rows than one tile holds run two strided passes each way (5 kernels per
    squaring); forced here on small sizes so the emulator covers the indexing and twiddles."""
    e = EmuEngine(p, 4, force_S, force_R2)
    assert e.R2 == force_R2
    wd = mo.widths(e.n, p)
    m = mr.mp(p)
    rng = random.Random(p + force_R2)
    x, y = rng.randrange(m), rng.randrange(m)
    put(e, 0, x, wd); put(e, 1, y, wd)
    for a in (1, 3, 1, 1):                       # a chain: the fused inverse+forward path kicks in from the 2nd
        e.square_mul(0, a); x = x * x * a % m
        assert val(e, 0) == x
    e.set_multiplicand(2, 1); e.mul(0, 2); x = x * y % m
    assert val(e, 0) == x
    e.sub(0, 2); e.square_mul(0); x = (x - 2) ** 2 % m
    assert val(e, 0) == x


def test_planner_decompositions():
    """The measured splits (marin_plan.hpp: measured_row_length, the radix-4 rule of the latency-bound sizes, rows of 8192 /
    10240 pairs for n = 2^24 / 5 * 2^22): which kernels a size runs on is decided here, so the table is pinned."""
    import ctypes as C
    L = lib()
    L.emu_plan.restype = C.c_int; L.emu_plan.argtypes = [C.c_uint32, C.POINTER(C.c_uint32)]
    def plan(p):
        out = (C.c_uint32 * 22)()
        assert L.emu_plan(p, out)
        return {"small": out[0], "S": out[1], "R": out[2], "R2": out[3], "stS": [out[6 + i] for i in range(out[4])], "stR": [out[14 + i] for i in range(out[5])]}
    want = {
        9941: dict(small=1),
        86243: dict(small=0, S=256, R=8, R2=1, stS=[4, 4, 4, 4], stR=[4, 2]),                      # n = 4096: radix-4 lists
        172001: dict(S=1024, R=4, stS=[4, 16, 16]),                                                # n = 8192: a specialised row shape
        204973: dict(S=64, R=80, stS=[4, 4, 4], stR=[5, 4, 4]),                                    # n = 10240
        1362763: dict(S=1024, R=32, stS=[4, 16, 16], stR=[4, 4, 2]),                               # n = 65536: strided pass on radix 4
        4999999: dict(S=2048, R=64, stS=[2, 4, 16, 16], stR=[4, 4, 4]),                            # n = 2^18: last size of the rule
        7240481: dict(S=1024, R=256, stS=[4, 16, 16], stR=[16, 16]),                               # n = 2^19
        136279841: dict(S=4096, R=1024, R2=1, stS=[16, 16, 16], stR=[16, 16, 4]),
        205271257: dict(S=5120, R=1024, R2=1, stS=[5, 4, 16, 16], stR=[16, 16, 4]),
        300000007: dict(S=8192, R=1024, R2=1, stS=[2, 16, 16, 16], stR=[16, 16, 4]),               # n = 2^24 on three passes
        340000019: dict(S=10240, R=1024, R2=1, stS=[5, 2, 4, 16, 16], stR=[16, 16, 4]),            # n = 5 * 2^22
        600000001: dict(S=4096, R=1024, R2=4, stR=[16, 16, 4]),                                    # n = 2^25: five passes, 1024 x 4 rows
        700000001: dict(S=5120, R=1024, R2=4), 1000000007: dict(S=8192, R=1024, R2=4), 1400000023: dict(S=10240, R=1024, R2=4),
        2147483647: dict(S=10240, R=1024, R2=8, stR=[16, 16, 4]),                                  # n = 5 * 2^25: 1024 x 8 rows
        3700000001: dict(S=10240, R=1024, R2=16),                                                  # n = 5 * 2^26: 1024 x 16 rows
    }
    for p, w in want.items():
        got = plan(p)
        for k, v in w.items():
            assert got[k] == v, (p, k, got)


@pytest.mark.parametrize("p,force_S", [(127, 0), (9941, 0), (86243, 0), (86243, 64), (216091, 0)])
def test_add_multiplicand_is_the_multiplicand_of_the_sum(p, force_S):
    """The transform is linear: the element-wise sum of two multiplicands multiplies like the sum of their residues
    (mb200_add_multiplicand; the V-trace term of P-1 stage 2).  Inputs include all-ones digits (the largest sums)."""
    e = EmuEngine(p, 8, force_S)
    assert e.max_multiplier() >= 2
    m = mr.mp(p)
    rng = random.Random(p + 17)
    for a, b, x in [(rng.randrange(m), rng.randrange(m), rng.randrange(m)), (m - 1, m - 2, m - 3), (0, 5, rng.randrange(m))]:
        e.set_mpz(0, a); e.set_mpz(1, b); e.set_mpz(2, x)
        e.set_multiplicand(3, 0); e.set_multiplicand(4, 1)
        e.add_multiplicand(5, 3, 4)
        e.mul(2, 5)
        assert val(e, 2) == x * (a + b) % m
        e.add_multiplicand(3, 3, 4)                      # in place
        e.set_mpz(2, x); e.mul(2, 3)
        assert val(e, 2) == x * (a + b) % m
