"""Generates tests/golden/pm1_stage1_res64.json: the P-1 stage-1 state X = 3^E mod 2^p - 1 (E = 2 p prod q^floor(log_q B1),
"fast3" chain of RunPM1.cpp:7127) for BASELINE config 4's exponent at a reduced B1, computed with GMP big integers
(oracle/_gmp.py).  The reference's own resume file for this exponent is a missing large blob.
Run from the repo root:  python tools/make_golden_pm1.py   (several minutes)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import _gmp                           # noqa: E402
from prmers_b200 import modes                     # noqa: E402

CASES = [(205271257, 100), (136279841, 64)]


def main():
    out = []
    for p, b1 in CASES:
        E = modes.pm1_exponent(p, b1)
        bits = bin(E)[2:]
        x = _gmp.mersenne_chain(p, 1, len(bits), mul_small=[3 if b == "1" else 1 for b in bits])
        out.append({"p": p, "b1": b1, "bits": len(bits), "res64": f"{x & 0xFFFFFFFFFFFFFFFF:016X}",
                    "res2048_low": f"{x & ((1 << 256) - 1):064x}"})
        print(out[-1], flush=True)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "pm1_stage1_res64.json")
    with open(path, "w") as f:
        json.dump({"what": "P-1 stage-1 state 3^E mod 2^p - 1, low 64 / 256 bits, GMP", "cases": out}, f, indent=1)


if __name__ == "__main__":
    main()
