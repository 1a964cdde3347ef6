"""The ECM driver's seed -> sigma -> twisted-Edwards curve chain (prmers_b200/ecm.py) against the REFERENCE'S OWN CODE:
oracle/_ref/libecm_ref.so is compiled from line ranges of /root/reference/src/modes/RunEcmTwistedEdwards.cpp where it
lies (oracle/Makefile, oracle/ref_ecm_shim.cpp).  Runs wherever the prebuilt shim and a libgmp runtime exist (this
container; the GPU box receives the built .so)."""
import ctypes as C
import os

import pytest

from prmers_b200 import ecm

_SO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libecm_ref.so")


def _lib():
    if not os.path.exists(_SO):
        pytest.skip("oracle/_ref/libecm_ref.so not built (needs the reference tree and libgmp: `make -C oracle ref`)")
    try:
        L = C.CDLL(_SO)
    except OSError as e:                                     # libgmp runtime absent
        pytest.skip(f"cannot load the reference ECM shim: {e}")
    L.ref_ecm_mix64.restype = C.c_uint64
    L.ref_ecm_mix64.argtypes = [C.c_uint64, C.c_uint64]
    L.ref_ecm_rnd_mpz_bits.argtypes = [C.c_uint32, C.c_uint64, C.c_uint, C.c_char_p, C.c_size_t]
    L.ref_ecm_curve_from_seed.restype = C.c_int
    L.ref_ecm_curve_from_seed.argtypes = [C.c_uint32, C.c_uint64] + [C.c_char_p] * 6 + [C.c_size_t]
    return L


def test_seed_mixing_matches_reference_code():
    L = _lib()
    for seed in (0, 1, 123456789, 0xDEADBEEFCAFEF00D, (1 << 64) - 1):
        for idx in (0, 1, 7, 255, 0x5EED5EED + 3, (1 << 63) + 5):
            assert ecm.mix64(seed, idx) == L.ref_ecm_mix64(seed, idx)
    buf = C.create_string_buffer(4096)
    for p in (127, 701, 1277):
        N = (1 << p) - 1
        for seed in (0, 5, 123456789, (1 << 64) - 3):
            for bits in (63, 64, 256, 1024):
                L.ref_ecm_rnd_mpz_bits(p, seed, bits, buf, len(buf))
                assert ecm.rnd_mpz_bits(N, seed, bits) == int(buf.value, 16), (p, seed, bits)


def test_curve_from_seed_matches_reference_code_1000_seeds():
    """Base seed -> curve seed (mix64) -> sigma -> (a, d, X0, Y0): every number equal to what the reference's own
    code computes, for 1000 curves on M701 (config 5) and a few on other moduli, including moduli with small factors
    where an inversion exposes one."""
    L = _lib()
    cap = 8192
    bufs = [C.create_string_buffer(cap) for _ in range(6)]
    checked = found = 0
    for p, base, count in ((701, 123456789, 1000), (1277, 42, 60), (101, 7, 60), (67, 3, 200), (59, 1, 200), (11, 9, 50)):
        N = (1 << p) - 1
        for c in range(count):
            cs = ecm.mix64(base, c)
            st = L.ref_ecm_curve_from_seed(p, cs, *bufs, cap)
            vals = [int(b.value or b"0", 16) for b in bufs]
            try:
                mine = ecm.curve_from_seed(N, cs)
            except ecm.FactorFound as e:
                assert st == 1 and vals[5] == e.g, (p, c)
                found += 1
                continue
            if mine is None:
                assert st == -1, (p, c)
                continue
            assert st == 0, (p, c)
            assert (mine["a"], mine["d"], mine["X0"], mine["Y0"], mine["sigma"]) == tuple(vals[:5]), (p, c)
            checked += 1
    assert checked >= 1300 and found >= 1
