"""Big-integer statements of what the modes compute (TEST INFRASTRUCTURE ONLY).

Independent of the IBDWT: plain Python integers mod 2^p - 1.  Used to pin the oracle (and the
CUDA engine) against the reference's published known-answer vectors.
Reference loops restated: RunPrpOrLlMarin.cpp:228-235,293-462 (PRP type 1 base 3, residue /9),
RunLlSafeMarin.cpp:481-507 (LL), RunPM1.cpp:7016-7135 (P-1 stage 1, E = 2 p prod q^floor(log_q B1)),
AlgoUtils.hpp:165-224 (res64 / res2048 formatting).
"""
from math import gcd

from . import _gmp


def mp(p):
    return (1 << p) - 1


def prp_residue(p, iters=None, base=3):
    """base^(2^iters) mod Mp (iters defaults to p, the type-1 PRP)."""
    return _gmp.mersenne_chain(p, base, p if iters is None else iters)


def prp_type1_res(p):
    """(is_prp, residue/9 mod Mp) as the reference reports it (RunPrpOrLlMarin.cpp:453-462)."""
    m = mp(p)
    x = prp_residue(p)
    r = x * pow(9, -1, m) % m if m % 3 != 0 else x
    return x == 9 % m, r


def res64_hex(r):
    return f"{r & 0xFFFFFFFFFFFFFFFF:016X}"


def res2048_hex(r):
    return f"{r & ((1 << 2048) - 1):0512x}"


def ll_residue(p, iters=None):
    return _gmp.mersenne_chain(p, 4, p - 2 if iters is None else iters, sub_small=2)


def primes_upto(n):
    sieve = bytearray([1]) * (n + 1)
    sieve[:2] = b"\0\0"
    for i in range(2, int(n ** 0.5) + 1):
        if sieve[i]:
            sieve[i * i::i] = bytearray(len(sieve[i * i::i]))
    return [i for i in range(n + 1) if sieve[i]]


def pm1_exponent(p, b1):
    e = 2 * p
    for q in primes_upto(b1):
        qk = q
        while qk * q <= b1:
            qk *= q
        e *= qk
    return e


def pm1_stage1(p, b1, base=3):
    """gcd(3^E - 1, Mp); returns (factor or None, X = 3^E mod Mp)."""
    m = mp(p)
    x = pow(base, pm1_exponent(p, b1), m)
    g = gcd((x - 1) % m, m)
    return (g if 1 < g < m else None), x


def pm1_stage2(p, b1, b2, base=3):
    """Textbook stage 2: gcd(prod (H^q - 1), Mp) over primes b1 < q <= b2."""
    m = mp(p)
    _, h = pm1_stage1(p, b1, base)
    acc = 1
    for q in primes_upto(b2):
        if q > b1:
            acc = acc * (pow(h, q, m) - 1) % m
    g = gcd(acc, m)
    return g if 1 < g < m else None


def mulmod(x, y, p):
    """x * y mod 2^p - 1 (GMP-backed for large p)."""
    return _gmp.mulmod_mersenne(x, y, p)
