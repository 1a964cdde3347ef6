"""Generates tests/golden/ecm_*.json: the ECM driver (prmers_b200/ecm.py) run on the big-integer engine
(oracle/int_engine.py), i.e. with exact arithmetic in place of every engine op.  The CUDA engine must
reproduce these residues and factors from the same seeds.  Every factor written is checked to divide 2^p - 1
(a proper divisor is self-certifying); the residues T after each stage are stored as SHA-256 prefixes.
The reference holds no (seed -> factor) vector for ECM, so these fixtures pin the driver to the
arithmetic, not to a run of the reference binary.

Run from the repo root:  python tools/make_golden_ecm.py
"""
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.int_engine import IntEngine          # noqa: E402
from prmers_b200 import ecm                      # noqa: E402

CASES = [  # (name, p, B1, B2, curves, seed)   -- config 5 of BASELINE.json first
    ("ecm_m701", 701, 6000, 33333, 8, 123456789),
    ("ecm_m101", 101, 300, 2000, 3, 123456789),
    ("ecm_m1277", 1277, 300, 1500, 2, 42),
]


def h(x):
    return None if x is None else hashlib.sha256(hex(x).encode()).hexdigest()[:32]


def main():
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, p, b1, b2, k, seed in CASES:
        r = ecm.run_ecm(lambda pp, regs: IntEngine(pp, regs), p, b1, b2, curves=k, seed=seed, continue_after_factor=True)
        N = (1 << p) - 1
        doc = {"p": p, "b1": b1, "b2": b2, "curves": k, "seed": seed, "k_bits": r["k_bits"], "status": r["status"],
               "factor": None if r["factor"] is None else str(r["factor"]), "per_curve": []}
        for c in r["curves"]:
            if c["factor"] is not None:
                assert 1 < c["factor"] < N and N % c["factor"] == 0
            doc["per_curve"].append({"curve": c["curve"], "curve_seed": str(c["curve_seed"]), "sigma": str(c.get("sigma")),
                                     "status": c["status"], "stage": c["stage"],
                                     "factor": None if c["factor"] is None else str(c["factor"]),
                                     "point_ops": c.get("point_ops"),
                                     "stage1_T_sha": h(c.get("stage1_T")), "stage2_T_sha": h(c.get("stage2_T"))})
        with open(os.path.join(out_dir, name + ".json"), "w") as f:
            json.dump(doc, f, indent=1)
        print(name, doc["status"], doc["factor"], [(c["status"], c["stage"], c["factor"]) for c in doc["per_curve"]])


if __name__ == "__main__":
    main()
