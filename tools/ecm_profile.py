"""Where the time of a batched ECM run goes (config 5, M701): host work per curve against the lockstep launches.
Usage: python tools/ecm_profile.py [curves] (GPU).  Prints the wall time of run_ecm_batched, the cProfile top list and the
time the stream needs for the point operations alone (the same scalar multiplication replayed with no host work)."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import prmers_b200 as pb
from prmers_b200 import ecm

curves = int(sys.argv[1]) if len(sys.argv) > 1 else 296
p, b1, b2, seed = 701, 6000, 33333, 123456789
bfactory = lambda q, regs, b: pb.Engine(q, regs, device=0, batch=b)
ecm.run_ecm_batched(bfactory, p, b1, b2, curves=8, seed=seed)       # warm-up (tables, programs)
for B in (8, curves):
    t0 = time.perf_counter()
    r = ecm.run_ecm_batched(bfactory, p, b1, b2, curves=B, seed=seed)
    dt = time.perf_counter() - t0
    print(f"B = {B}: {dt:.3f} s, {B / dt:.1f} curves/s, found {sum(1 for c in r['curves'] if c['status'] == 'found')}")
    # the point operations alone: stage-1 scalar on the batched engine, stream time by events
    eng = bfactory(p, ecm.REGS, B)
    for rg in range(ecm.REGS):
        eng.set(rg, 3 + rg)
    ops = ecm.EdwardsOps(eng)
    digits = ecm.naf(ecm.stage1_exponent(b1))
    eng.sync()
    t0 = time.perf_counter()
    ecm._scalar_mul(eng, ops, digits)
    t_enq = time.perf_counter() - t0
    eng.sync()
    t_all = time.perf_counter() - t0
    print(f"B = {B}: stage-1 scalar alone: {ops.ops} point ops, enqueue {t_enq:.3f} s, done {t_all:.3f} s, {t_all * 1e6 / max(ops.ops, 1):.1f} us per point op")
    eng.close()
pr = cProfile.Profile()
pr.enable()
ecm.run_ecm_batched(bfactory, p, b1, b2, curves=curves, seed=seed)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
