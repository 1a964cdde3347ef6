"""Development aid: squaring time per row length S of the contiguous pass (MARIN_B200_FORCE_S), for the planner's table."""
import os, subprocess, sys
here = os.path.dirname(os.path.abspath(__file__))
cases = [int(a) for a in sys.argv[1:]] or [2500009, 4999999, 9999991, 20996011, 37156667, 77232917]
for p in cases:
    for S in (0, 512, 1024, 2048, 4096, 5120):
        env = dict(os.environ)
        if S: env["MARIN_B200_FORCE_S"] = str(S)
        try:
            r = subprocess.run([sys.executable, os.path.join(here, "pass_times.py"), str(p)], env=env, capture_output=True, text=True, timeout=120)
            out = (r.stdout.strip().splitlines() or [r.stderr.strip()[-200:]])[-1]
        except subprocess.TimeoutExpired:
            out = "timeout"
        print(f"S={S:5d} {out}", flush=True)
