"""Development aid: squaring time at the latency-bound sizes per row length S and radix-4 mask (1 = contiguous pass on
radix-4 stages, 2 = strided pass).  The mask was a build-time experiment (MARIN_B200_R4MASK read by make_plan); its outcome
is the planner's r4S / r4R rule, the log is profiles/r2/lowlat_sweep.log."""
import os, subprocess, sys
here = os.path.dirname(os.path.abspath(__file__))
cases = [int(a) for a in sys.argv[1:]] or [86243, 172001, 350003, 700001, 1362763, 2500009, 4999999]
for p in cases:
    for S in (0, 64, 128, 256, 512, 1024, 2048):
        for mask in (0, 1, 2, 3):
            env = dict(os.environ)
            if S: env["MARIN_B200_FORCE_S"] = str(S)
            if mask: env["MARIN_B200_R4MASK"] = str(mask)
            try:
                r = subprocess.run([sys.executable, os.path.join(here, "pass_times.py"), str(p)], env=env, capture_output=True, text=True, timeout=120)
                out = (r.stdout.strip().splitlines() or [r.stderr.strip()[-120:]])[-1]
            except subprocess.TimeoutExpired:
                out = "timeout"
            print(f"S={S:5d} mask={mask} {out}", flush=True)
