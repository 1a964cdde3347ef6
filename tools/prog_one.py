"""One recorded Edwards doubling at M701 replayed a few times (the ncu target for k_program)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import prmers_b200 as pb
from prmers_b200 import ecm

e = pb.Engine(int(sys.argv[1]) if len(sys.argv) > 1 else 701, ecm.REGS, device=0)
for r in range(ecm.REGS):
    e.set(r, 3 + r)
e.set_multiplicand(43, 16)
ops = ecm.EdwardsOps(e, graphs=True)
for _ in range(60):
    ops.dbl()
e.sync()
print("ok", e.get_mpz(3) % 1000003)
