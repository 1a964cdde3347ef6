"""Development aid: inverse strided pass with parts left out (library built with -DMARIN_DEBUG_SKIP; results are wrong)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os
sys.path.insert(0, %r)
import prmers_b200 as pb
e = pb.Engine(136279841, 3)
e.set(0, 3); e.square_mul_repeat(0, 20); e.sync()
print(" ".join("%%.1f" %% (e.bench_pass(1, w, 100) / 100 * 1000) for w in (2, 3, 4)))
''' % ROOT
names = {1: "no gate wait", 2: "no reweight mul", 3: "no unweight mul+loads", 4: "no carry_run", 5: "no reweight/store phase",
         6: "no carry phase", 7: "no slot store", 8: "no absorb phase"}
names[9] = 'no dst store'
for mask in (0, 512, 32):
    env = dict(os.environ, MARIN_B200_LIB=os.path.join(ROOT, "prmers_b200", "alt_dbg.so"), MARIN_B200_DBG=str(mask))
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=120)
    print(f"mask {mask:2d} [{', '.join(n for b, n in names.items() if mask >> b & 1) or 'full'}]: inv+carry / inv+carry+fwd / inv-plain = {r.stdout.strip() or r.stderr.strip()[-300:]}", flush=True)
