"""What a recorded program costs inside k_program (GPU): launch-to-launch time of programs of different shapes at the ECM
size (M701, n = 32), so that the per-launch cost, the cost of an op in a chain and the gain of the warp groups separate.
Usage: python tools/prog_probe.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import prmers_b200 as pb
from prmers_b200 import ecm

p = int(sys.argv[1]) if len(sys.argv) > 1 else 701
e = pb.Engine(p, ecm.REGS, device=0)
for r in range(ecm.REGS):
    e.set(r, 3 + r)
e.set_multiplicand(43, 16); e.set_multiplicand(46, 9); e.set_multiplicand(45, 29); e.set_multiplicand(50, 49)
ops = ecm.EdwardsOps(e, graphs=False)


def timed(name, body, reps=3000):
    e.graph_begin()
    body()
    g = e.graph_end()
    for _ in range(200):
        e.graph_launch(g)
    e.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        e.graph_launch(g)
    e.sync()
    us = (time.perf_counter() - t0) * 1e6 / reps
    print(f"{name:58s} {us:8.2f} us per launch")
    return us


def rep(fn, k):
    def body():
        for _ in range(k):
            fn()
    return body


t1 = timed("1 x copy (the launch itself)", lambda: e.copy(20, 21))
timed("8 x copy in a chain (20 -> 21 -> 20 ...)", lambda: [e.copy(20 + (i & 1), 21 - (i & 1)) for i in range(8)])
timed("8 x independent copy", lambda: [e.copy(20 + i, 30 + i) for i in range(8)])
timed("1 x square_mul", lambda: e.square_mul(20))
timed("8 x square_mul in a chain", rep(lambda: e.square_mul(20), 8))
timed("8 x square_mul on 8 registers", lambda: [e.square_mul(20 + i) for i in range(8)])
timed("4 x square_mul on 4 registers", lambda: [e.square_mul(20 + i) for i in range(4)])
timed("8 x add in a chain", rep(lambda: e.add(20, 21), 8))
timed("8 x set_multiplicand on 8 registers", lambda: [e.set_multiplicand(30 + i, 20 + i) for i in range(8)])
d1 = timed("1 x Edwards doubling (23 ops)", lambda: ops._dbl())
d4 = timed("4 x Edwards doubling", rep(lambda: ops._dbl(), 4), reps=1500)
print(f"  -> {(d4 - d1) / 3:.2f} us per doubling inside the kernel, {d1 - (d4 - d1) / 3:.2f} us per launch around it")
a1 = timed("1 x Edwards mixed addition (40 ops)", lambda: ops._add(False))
a4 = timed("4 x Edwards mixed addition", rep(lambda: ops._add(False), 4), reps=1500)
print(f"  -> {(a4 - a1) / 3:.2f} us per addition inside the kernel")
