"""Development aid: cycles per phase of k_mid for a few CTAs (needs a library built with -DMARIN_PHASE_TIMING)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import prmers_b200 as pb
L = pb.load_library()
e = pb.Engine(136279841, 4)
e.set(0, 3); e.square_mul_repeat(0, 50); e.sync()
e.bench_pass(1, 1, 3)
buf = (C.c_longlong * 128)()
L.mb200_debug_phase_clocks.argtypes = [C.c_void_p]
L.mb200_debug_phase_clocks(buf)
for s in range(4):
    c = [buf[s * 32 + i] for i in range(9)]
    print("block slot", s, "phase cycles:", [c[i + 1] - c[i] for i in range(7)], "total", c[7] - c[0])
