"""Generates tests/golden/pm1_stage2_vtrace.json: the V-trace stage-2 accumulator of BASELINE config 4's exponent
(p = 205271257, n = 5 * 2^21) at reduced bounds, computed WITHOUT the product's driver: V_n by the plain three-term
recurrence V_{n+1} = V_1 V_n - V_{n-1} on GMP integers (oracle/_gmp.py) for every n up to the last giant step, then
the product of (V_{kD} - V_j) over the (k, j) with a prime kD - j or kD + j in (B1, B2].  V_1 = 3^(2^40) mod 2^p - 1,
which the GPU test forms with 40 squarings (no host-side inverse of a 2 * 10^8-bit number in a test).
Run from the repo root:  python tools/make_golden_pm1_stage2.py   (about ten minutes)."""
import hashlib
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import _gmp                           # noqa: E402

CASES = [dict(p=205271257, v1_squarings=40, b1=100, b2=400, D=30), dict(p=1362763, v1_squarings=40, b1=100, b2=3000, D=210)]


def is_prime(n):
    return n > 1 and all(n % q for q in range(2, int(n ** 0.5) + 1))


def main():
    out = []
    for c in CASES:
        p, b1, b2, D = c["p"], c["b1"], c["b2"], c["D"]
        m = (1 << p) - 1
        v1 = _gmp.mersenne_chain(p, 3, c["v1_squarings"])
        k_first, k_last = (b1 + 1 + D // 2) // D, (b2 + D // 2) // D
        V = [2, v1]
        for n in range(1, k_last * D + 1):
            V.append((_gmp.mulmod_mersenne(v1, V[n], p) - V[n - 1]) % m)
            if n % 50 == 0:
                print(f"p={p}: V_{n}", flush=True)
        babies = [j for j in range(1, D // 2 + 1) if math.gcd(j, D) == 1]
        acc, terms, primes = 1, 0, 0
        for k in range(k_first, k_last + 1):
            for j in babies:
                lo, hi = k * D - j, k * D + j
                a, b = b1 < lo <= b2 and is_prime(lo), b1 < hi <= b2 and is_prime(hi)
                if a or b:
                    acc = _gmp.mulmod_mersenne(acc, (V[k * D] - V[j]) % m, p)
                    terms += 1; primes += int(a) + int(b)
        raw = acc.to_bytes((p + 7) // 8, "little")
        out.append(dict(c, terms=terms, primes=primes, res64=f"{acc & 0xFFFFFFFFFFFFFFFF:016X}", sha256=hashlib.sha256(raw).hexdigest()))
        print(out[-1], flush=True)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "pm1_stage2_vtrace.json")
    with open(path, "w") as f:
        json.dump({"what": "V-trace stage-2 accumulator prod (V_kD - V_j), V_1 = 3^(2^s), plain recurrence on GMP integers", "cases": out}, f, indent=1)


if __name__ == "__main__":
    main()
