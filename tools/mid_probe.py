"""Development aid: k_mid time versus grid size and CTA width (lone-CTA latency, wave quantisation)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os
sys.path.insert(0, %r)
import prmers_b200 as pb
e = pb.Engine(136279841, 3)
e.set(0, 3); e.square_mul_repeat(0, 20); e.sync()
print("%%.1f" %% (e.bench_pass(1, 1, 100) / 100 * 1000))
''' % ROOT
def run(lib, grid, thr):
    env = dict(os.environ)
    if lib: env["MARIN_B200_LIB"] = os.path.join(ROOT, "prmers_b200", lib)
    if grid: env["MARIN_B200_MID_GRID"] = str(grid)
    if thr: env["MARIN_B200_MID_THREADS"] = str(thr)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=120)
    return r.stdout.strip() or r.stderr.strip()[-200:]
for lib, thr in (("", 0), ("alt_mid768.so", 256), ("alt_mid768.so", 512), ("alt_mid768.so", 768)):
    print(lib or "default", "threads", thr or 256, {g: run(lib, g, thr) for g in (148, 296, 444, 592, 888, 1024)}, flush=True)
