"""Per-operation timings of the engine API (development aid / DESIGN.md table): wall clock around
`count` asynchronous calls + sync, through the C ABI from Python."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import prmers_b200 as pb

def t(e, f, count):
    f(); e.sync()
    t0 = time.perf_counter()
    for _ in range(count): f()
    e.sync()
    return (time.perf_counter() - t0) / count * 1e6

CASES = [(205271257, 14), (136279841, 8), (1362763, 8), (86243, 8), (701, 51), (127, 18)]
if len(sys.argv) > 1:
    CASES = [(int(a), 8) for a in sys.argv[1:]]
for p, regs in CASES:
    e = pb.Engine(p, regs)
    e.set(0, 3); e.set(1, 5); e.set(3, 7)
    e.square_mul_repeat(0, 20); e.copy(1, 0); e.copy(3, 0); e.set_multiplicand(2, 1)
    cnt = 200 if e.get_size() >= (1 << 20) else 3000
    a = 3 if e.max_multiplier() >= 3 else 1
    r = {
        "square_mul": t(e, lambda: e.square_mul(0), cnt),
        f"square_mul(a={a})": t(e, lambda: e.square_mul(0, a), cnt),
        "mul": t(e, lambda: e.mul(3, 2), cnt),
        "set_multiplicand": t(e, lambda: e.set_multiplicand(4, 1), cnt),
        "add": t(e, lambda: e.add(3, 1), cnt),
        "sub_reg": t(e, lambda: e.sub_reg(3, 1), cnt),
        "sub(2)": t(e, lambda: e.sub(3, 2), cnt),
        "copy": t(e, lambda: e.copy(5, 1), cnt),
        "addsub": t(e, lambda: e.addsub(5, 6, 1, 3), cnt),
        "addsub_copy": t(e, lambda: e.addsub_copy(5, 6, 4, 7, 1, 3), cnt),
        "square_mul_copy": t(e, lambda: e.square_mul_copy(0, 5), cnt),
        "mul_copy": t(e, lambda: e.mul_copy(3, 2, 5), cnt),
        "mul_add": t(e, lambda: e.mul_add(3, 2, 1), cnt),
        "is_equal": t(e, lambda: e.is_equal(0, 1), 20),
        "get": t(e, lambda: e.get(0), 10),
    }
    print(f"p={p} n={e.get_size()}: " + "  ".join(f"{k} {v:.1f}" for k, v in r.items()) + "  (us/op)", flush=True)
    e.close()
