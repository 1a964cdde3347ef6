"""Development aid: per-pass and total times for one exponent."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import prmers_b200 as pb
p = int(sys.argv[1]) if len(sys.argv) > 1 else 136279841
e = pb.Engine(p, 4, force_S=int(os.environ.get('MARIN_B200_FORCE_S', 0)))  # the sweep tools pass the row length this way
e.set(0, 3); e.square_mul_repeat(0, 100); e.sync()
ms = e.bench_square_mul(0, 1000) / 1000
ps = [e.bench_pass(1, w, 200) / 200 for w in (0, 1, 2, 4, 3)]
print(f"p={p} n={e.get_size()} total {ms*1000:.1f} us  fwd {ps[0]*1000:.1f}  mid {ps[1]*1000:.1f}  inv {ps[2]*1000:.1f}  inv-plain {ps[3]*1000:.1f}  inv+carry+fwd {ps[4]*1000:.1f}", flush=True)

