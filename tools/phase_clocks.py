"""Development aid: SM cycles per phase of k_inv_fast_T4<true> for every warp of four tiles in the middle of the grid
(needs a library built with -DMARIN_PHASE_TIMING; MARIN_B200_LIB points at it)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import prmers_b200 as pb
L = pb.load_library()
e = pb.Engine(136279841, 4)
e.set(0, 3); e.square_mul_repeat(0, 50); e.sync()
buf = (C.c_longlong * (4 * 8 * 32))()
L.mb200_debug_phase_clocks.argtypes = [C.c_void_p]
L.mb200_debug_phase_clocks(buf)
names = {0: "entry", 1: "ticket+setup", 2: "i.ph0", 3: "-", 4: "i.ph0 done", 5: "i.ph1", 6: "-", 7: "i.ph1 done", 8: "i.ph2 (at barrier)", 9: "barrier passed",
         10: "i.ph2 bfly done", 11: "publish done/i.ph3", 12: "-", 13: "i.ph3 bfly done", 14: "publish done", 15: "absorb0 done", 16: "absorb1 done",
         17: "barrier passed", 18: "f.ph0", 19: "-", 20: "f.ph1", 21: "-", 22: "f.ph2 (at barrier)", 23: "barrier passed", 24: "f.ph3", 25: "-", 26: "end"}
for s in range(4):
    t0 = min(buf[(s * 8 + w) * 32] for w in range(8))
    print(f"tile slot {s}")
    for w in range(8):
        c = [buf[(s * 8 + w) * 32 + i] - t0 for i in range(27)]
        print(f"  warp {w}: " + " ".join(f"{i}:{c[i]}" for i in range(27) if names[i] != "-"))
print("stamps:", {i: n for i, n in names.items() if n != "-"})
