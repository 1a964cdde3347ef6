"""Config 5 timing (development aid / DESIGN.md table): one ECM curve on M701 (n = 32) with and without CUDA-graph
replay of the point operations; reports microseconds per point operation, per engine op, and curves per second."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import prmers_b200 as pb
from prmers_b200 import ecm

p, b1, b2 = 1277, 6000, 33333           # M1277 has no small factor: stage 1 and stage 2 run to the end at the M701-class size
if len(sys.argv) > 1: p = int(sys.argv[1])
for graphs in (True, False):
    e = pb.Engine(p, ecm.REGS)
    l0 = e.launch_count()
    t0 = time.perf_counter()
    r = ecm.run_curve(e, p, b1, b2, ecm.mix64(123456789, 0), graphs=graphs)
    e.sync()
    dt = time.perf_counter() - t0
    nl = e.launch_count() - l0
    print(f"p={p} n={e.get_size()} graphs={graphs}: status {r['status']} stage {r['stage']}  {dt:.2f} s/curve  {r['point_ops']} point ops  "
          f"{dt / r['point_ops'] * 1e6:.1f} us/point-op  {nl} launches  {dt / nl * 1e6:.2f} us/launch", flush=True)
    e.close()
