"""Generates tests/golden/prp_interim_res64.json: res64 of 3^(2^i) mod 2^p - 1 at a few iteration counts for the
BASELINE.json exponents, computed with GMP big integers (oracle/_gmp.py: plain squaring + Mersenne folding, no
transform at all).  The reference stores no interim residues for these exponents, so this pins the
engine at full size to independent exact arithmetic.  Run from the repo root:  python tools/make_golden_prp.py
(about ten minutes: a 136-Mbit GMP squaring takes a second or two)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import _gmp                           # noqa: E402

CASES = {136279841: [1, 27, 28, 40, 100, 200], 205271257: [1, 28, 29, 60, 120], 139123459: [50]}


def main():
    assert _gmp.available(), "needs the libgmp runtime"
    out = {}
    for p, marks in CASES.items():
        x, done, res, t0 = 3, 0, {}, time.time()
        for it in marks:
            x = _gmp.mersenne_chain(p, x, it - done)
            done = it
            res[str(it)] = f"{x & 0xFFFFFFFFFFFFFFFF:016X}"
            print(p, it, res[str(it)], f"{time.time() - t0:.0f}s", flush=True)
        out[str(p)] = res
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "prp_interim_res64.json")
    with open(path, "w") as f:
        json.dump({"what": "low 64 bits of 3^(2^i) mod 2^p - 1 (type-1 PRP state after i squarings), GMP", "res64": out}, f, indent=1)


if __name__ == "__main__":
    main()
