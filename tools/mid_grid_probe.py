import os, subprocess, sys
ROOT = "/root/repo"
CHILD = r'''
import sys
sys.path.insert(0, %r)
import prmers_b200 as pb
e = pb.Engine(136279841, 3)
e.set(0, 3); e.square_mul_repeat(0, 20); e.sync()
print("%%.1f" %% (e.bench_pass(1, 1, 200) / 200 * 1000))
''' % ROOT
for g in (148, 296, 444, 592, 740, 888, 1024, 1036):
    env = dict(os.environ); env["MARIN_B200_MID_GRID"] = str(g)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=120)
    print(g, r.stdout.strip() or r.stderr.strip()[-300:], flush=True)
