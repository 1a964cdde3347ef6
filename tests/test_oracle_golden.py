"""Pins the CPU oracle (oracle/marin_oracle.c): (1) against the reference's own headers compiled
in place (oracle/_ref, skipped where /root/reference is absent and no prebuilt shim travelled),
(2) against the reference's known-answer vectors (unit_tests.sh), (3) against big-int math."""
import ctypes as C
import random

import numpy as np
import pytest

from oracle import marin_oracle as mo
from oracle import mersenne_ref as mr

P = mo.MOD_P
EXPONENTS = [31, 89, 127, 521, 701, 1279, 9941, 11213, 86243, 100003, 216091, 1362763,
             57885161, 136279841, 161051017, 175000001, 205271257, 600000001]


def _ref():
    r = mo.ref_lib()
    if r is None:
        pytest.skip("oracle/_ref not built (reference tree absent)")
    return r


def test_field_ops_match_reference_headers():
    R, L = _ref(), mo.lib()
    rng = random.Random(1)
    edge = [0, 1, 2, P - 1, P - 2, 0xFFFFFFFF, 0x100000000, 0xFFFFFFFF00000000, (1 << 63), 7, 554]
    vals = edge + [rng.randrange(P) for _ in range(400)]
    for a in vals:
        assert L.mo_mod_inv(a or 1) == R.ref_mod_invert(a or 1)
        for b in vals[:40]:
            assert L.mo_mod_add(a, b) == R.ref_mod_add(a, b) == (a + b) % P
            assert L.mo_mod_sub(a, b) == R.ref_mod_sub(a, b) == (a - b) % P
            assert L.mo_mod_mul(a, b) == R.ref_mod_mul(a, b) == (a * b) % P
    assert R.ref_mod_muli(1) == pow(2, 48, P) and pow(2, 96, P) == P - 1
    assert pow(554, (P - 1) // 192, P) == 2          # n-th roots of two exist (ibdwt.h:116)
    for n in (4, 8, 40, 4096, 5 << 21, 1 << 23):
        assert L.mo_root_nth(n) == R.ref_mod_root_nth(n) == pow(7, (P - 1) // n, P)


def test_adc_helpers_match_reference_headers():
    R, L = _ref(), mo.lib()
    rng = random.Random(2)
    for _ in range(2000):
        lhs = rng.randrange(1 << 64) if rng.random() < 0.5 else rng.randrange(1 << 20)
        width = rng.randrange(1, 32)
        cin = rng.randrange(1 << 64) if rng.random() < 0.3 else rng.randrange(1 << 40)
        a = rng.randrange(1, 1 << 32) if rng.random() < 0.5 else 3
        c1, c2 = C.c_uint64(cin), C.c_uint64(cin)
        assert L.mo_adc(lhs, width, C.byref(c1)) == R.ref_adc(lhs, width, C.byref(c2))
        assert c1.value == c2.value
        c1, c2 = C.c_uint64(cin), C.c_uint64(cin)
        assert L.mo_adc_mul(lhs, a, width, C.byref(c1)) == R.ref_adc_mul(lhs, a, width, C.byref(c2))
        assert c1.value == c2.value


def test_planner_matches_reference_and_its_pinned_sizes():
    L = mo.lib()
    # tests/test_aevum_auto_policy.cpp:12-28 pins these Marin sizes
    assert L.mo_transform_size(1362763) == 65536
    assert L.mo_transform_size(136279841) == 8388608
    assert L.mo_transform_size(175000001) == 10485760
    # sizes at the BASELINE configs
    assert [L.mo_transform_size(p) for p in (127, 86243, 205271257, 701)] == [8, 4096, 5 << 21, 32]
    R = mo.ref_lib()
    if R is not None:
        rng = random.Random(3)
        ps = EXPONENTS + [rng.randrange(3, 1 << 31) for _ in range(3000)] + list(range(3, 600))
        for p in ps:
            assert L.mo_transform_size(p) == R.ref_transform_size(p), p


@pytest.mark.parametrize("p", [31, 127, 521, 701, 11213, 86243, 216091, 1362763])
def test_widths_weights_match_reference_headers(p):
    R = _ref()
    n = mo.transform_size(p)
    w, wi = mo.weights(n, p)
    wd = mo.widths(n, p)
    rw = np.zeros(2 * n, dtype=np.uint64)
    rwd = np.zeros(n, dtype=np.uint8)
    assert R.ref_weights_widths(n, p, rw.ctypes.data, rwd.ctypes.data) == 0
    assert np.array_equal(wd, rwd)
    assert int(wd.astype(np.int64).sum()) == p
    j = np.arange(n)
    idx = (j % 4) * (n // 4) + j // 4          # reference table layout, ibdwt.h:142
    assert np.array_equal(w, rw[2 * idx])
    assert np.array_equal(wi, rw[2 * idx + 1])


def test_canonical_export_matches_reference_get():
    """engine_gpu_flat::get (engine_gpu.h:1534-1561) restated by mo_get_digits."""
    R = _ref()
    for p in (127, 701, 11213):
        n = mo.transform_size(p)
        e = mo.OracleEngine(p, 2)
        e.set(0, 3)
        for _ in range(40):
            e.square_mul(0, 3)
        raw = np.zeros(n, dtype=np.uint64)
        mo.lib().mo_get_raw(e._h, 0, raw.ctypes.data)
        rw = np.zeros(2 * n, dtype=np.uint64)
        rwd = np.zeros(n, dtype=np.uint8)
        R.ref_weights_widths(n, p, rw.ctypes.data, rwd.ctypes.data)
        out = np.zeros(n, dtype=np.uint64)
        R.ref_get_canonical(n, rw.ctypes.data, rwd.ctypes.data, raw.ctypes.data, out.ctypes.data)
        assert np.array_equal(out, e.get_digits(0))


# ---- the reference's known-answer vectors (pure math, independent of the engine) ----

M11213_RES64 = {1000: "FBA631FBCB73A011", 2000: "F01283650C4A1491", 3000: "7E79193B757010B7",
                4000: "31482E4D80FE99BB", 5000: "973B76BACF73BBEF", 6000: "8CFFB332495FC320",
                7000: "98080C76DF068843", 8000: "8FDA516F885D3FEE", 9000: "2AADBC4F1E318E92",
                10000: "0A4AAF339C8B290C", 11000: "A1F26F470CFE412D"}   # unit_tests.sh:164-187


def test_m11213_interim_res64_oracle_engine():
    e = mo.OracleEngine(11213, 2)
    assert e.n == 512
    e.set(0, 3)
    for it in range(1, 11214):
        e.square_mul(0)
        if it in M11213_RES64:
            assert mr.res64_hex(mo.digits_to_int(e.get_digits(0))) == M11213_RES64[it], it
    assert e.get_int(0) == 9                                   # M11213 is prime (unit_tests.sh:151-161)


M100003_RES2048 = ("af262d00ed00a05d53e99d0e0e451b12405ddabe139fe8396a4c520b505bb65bed1609d3c8ef23bb"
                   "b1d0f8140a6bcdd2c67f9c8aa3bd0e6eeb3e8e79db904810c88de09820557176b389290f84f18424"
                   "efa6a59fb9f132a74f53a83ba6e2f508c617a5e1451c3ee08d179e6614026f973d1900602f2068a0"
                   "8894cd81ed5035de9ded85909b1ee6ff4dc723118b79d3f940272ae1066aebe27c86338ad7edf70e"
                   "76c0e8abf3e985b73db2a06f1b742a9a908728be2bd4b7daa2d6aafc11bacaaa40944e9a66b039cb"
                   "0deaaa8e5e357cd54b81b3ec6661d55e48bacb994bfd3cbb33f3f01d82347fa00578ec86c4cd7eb5"
                   "68a1463cf3e38dae1cf45e9503c71fd6")                  # unit_tests.sh:137-148


def test_m100003_res64_res2048_bigint():
    is_prp, r = mr.prp_type1_res(100003)
    assert not is_prp
    assert mr.res64_hex(r) == "1CF45E9503C71FD6"
    assert mr.res2048_hex(r) == M100003_RES2048


def test_m100003_oracle_engine_matches_bigint_prefix():
    p = 100003
    e = mo.OracleEngine(p, 1)
    e.set(0, 3)
    for _ in range(300):
        e.square_mul(0)
    assert e.get_int(0) == mr.prp_residue(p, 300)


PM1_STAGE1 = [(269, 2141, 13822297), (269, 192, None), (541, 899, 4312790327), (263, 3527, 23671),
              (367, 11981, 646300400639), (2151, 256, 327405968242246366421788399),
              (544139, 3, 22853839), (44159, 23, None), (144139, 3, None), (11544157, 19, None)]


@pytest.mark.parametrize("p,b1,factor", PM1_STAGE1[:8])
def test_pm1_stage1_vectors_bigint(p, b1, factor):
    assert mr.pm1_stage1(p, b1)[0] == factor               # unit_tests.sh:54-69,204-213


def test_pm1_stage2_vectors_bigint():
    assert mr.pm1_stage2(269, 4, 2141) == 13822297
    assert mr.pm1_stage2(569, 9, 677) == 55470673
    assert mr.pm1_stage2(1097, 3, 709) == 4576661533441
    assert mr.pm1_stage2(4133, 23, 2099) == 11173615097
    assert mr.pm1_stage2(269, 192, 457) is None


def test_pm1_stage1_on_oracle_engine():
    """fast3 loop of RunPM1.cpp:7075-7135: X <- X^2 * (3 if bit else 1), MSB first."""
    for p, b1, factor in [(269, 2141, 13822297), (541, 899, 4312790327), (367, 11981, 646300400639)]:
        E = mr.pm1_exponent(p, b1)
        e = mo.OracleEngine(p, 1)
        e.set(0, 1)
        for bit in bin(E)[2:]:
            e.square_mul(0, 3 if bit == "1" else 1)
        x = e.get_int(0)
        assert x == mr.pm1_stage1(p, b1)[1]
        from math import gcd
        assert gcd(x - 1, mr.mp(p)) == factor


@pytest.mark.parametrize("p", [89, 107, 127, 521, 607, 1279, 2203, 2281, 3217, 4253, 4423])
def test_small_mersenne_primes_prp_and_ll(p):
    e = mo.OracleEngine(p, 2)
    e.set(0, 3)
    for _ in range(p):
        e.square_mul(0)
    assert e.get_int(0) == 9                                   # unit_tests.sh:5-10
    e.set(1, 4)
    for _ in range(p - 2):
        e.square_mul(1)
        e.sub(1, 2)
    assert e.get_int(1) == 0                                   # LL: s_{p-2} = 0


@pytest.mark.parametrize("p", [57, 91, 100, 200, 500, 1001, 4095, 8191 * 0 + 8191])
def test_composites_fail_ll(p):
    if p == 8191:
        pytest.skip("8191 is prime; the reference lists it under -ll composites by mistake")
    e = mo.OracleEngine(p, 1)
    e.set(0, 4)
    for _ in range(p - 2):
        e.square_mul(0)
        e.sub(0, 2)
    assert e.get_int(0) == mr.ll_residue(p) != 0              # unit_tests.sh:12-14


@pytest.mark.parametrize("p,n", [(127, 8), (701, 32), (367, 16), (1279, 64), (521, 40), (216091, 10240),
                                 (9941, 512), (3217, 160)])
def test_engine_ops_vs_bigint(p, n):
    """Register API semantics (engine.h:45-74, tests/test_aevum_reg_adapter.cpp:27-87) incl. 5*2^k."""
    try:
        e = mo.OracleEngine(p, 6, n=n)
    except ValueError:
        pytest.skip("size not valid for exponent")
    m = mr.mp(p)
    rng = random.Random(p)
    x, y = rng.randrange(m), rng.randrange(m)
    e.set_int(0, x); e.set_int(1, y)
    assert e.get_int(0) == x and e.get_int(1) == y
    e.set_multiplicand(2, 1)
    e.mul(0, 2, 3); x = x * y * 3 % m
    assert e.get_int(0) == x
    e.square_mul(0, 7); x = x * x * 7 % m
    assert e.get_int(0) == x
    e.add(0, 1); x = (x + y) % m
    assert e.get_int(0) == x
    e.sub_reg(0, 1); e.sub_reg(0, 1); x = (x - 2 * y) % m
    assert e.get_int(0) == x
    e.sub(0, 2); x = (x - 2) % m
    assert e.get_int(0) == x
    e.add(0, 0); x = 2 * x % m                                  # aliasing add(R,R)
    assert e.get_int(0) == x
    e.copy(3, 0); e.set(4, 0); e.sub(4, 5)
    assert e.get_int(3) == x and e.get_int(4) == m - 5
    e.set(5, 0xFFFFFFFF)
    assert e.get_int(5) == 0xFFFFFFFF % m
