"""Development aid: the specialised kernels (marin_fast.cuh) against the generic ones, raw register images, bit for bit."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import prmers_b200 as pb

def run(p, fast):
    if fast: os.environ.pop("MARIN_B200_NO_FAST", None)
    else: os.environ["MARIN_B200_NO_FAST"] = "1"
    e = pb.Engine(p, 4)
    n = e.get_size()
    rng = np.random.default_rng(7)
    out = []
    def raw(r):
        b = bytearray(e.get_register_data_size()); e.get_data(b, r); return np.frombuffer(bytes(b), dtype=np.uint64)
    e.set(0, 3)
    for _ in range(40): e.square_mul(0)
    e.copy(1, 0)
    for _ in range(5): e.square_mul(0)
    out.append(raw(0))
    e.square_mul(0, 3)
    out.append(raw(0))
    e.set_multiplicand(2, 1)
    out.append(raw(2))
    e.mul(0, 2, 1)
    out.append(raw(0))
    e.mul(1, 2, 3)
    out.append(raw(1))
    e.square_mul(1); e.square_mul(1); e.square_mul(1)
    e.add(1, 0)
    out.append(raw(1))
    out.append(e.get(1))
    e.close()
    return out

ok = True
for p in [int(a) for a in sys.argv[1:]] or [136279841, 110503001]:
    a = run(p, True); b = run(p, False)
    for i, (x, y) in enumerate(zip(a, b)):
        same = np.array_equal(np.asarray(x), np.asarray(y))
        ok &= same
        print(f"p={p} step {i}: {'identical' if same else 'DIFFERENT'}", flush=True)
print("FAST_CHECK", "OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
