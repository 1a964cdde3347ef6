"""The reference's own benchmark ladder (`prmers -bench`, src/core/App.cpp:670-746): for each exponent a 6-register
engine, 96 warm-up squarings, then `square_mul(R0)` in a loop -- here timed with CUDA events AND a final sync (the
reference reads the wall clock without one).  One line per exponent: transform size, us per squaring, iterations/s."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import prmers_b200 as pb

EXPS = [127, 1279, 2203, 9941, 44497, 756839, 3021377, 37156667, 57885161, 77232917, 82589933, 136279841, 146410013, 161051017,
        177156127, 180000017, 200000033, 220000013, 250000013, 280000027, 300000007, 320000077, 340000019, 360000019, 400000009,
        500000003, 600000001]
rows = []
for p in [int(a) for a in sys.argv[1:]] or EXPS:
    e = pb.Engine(p, 6)
    e.set(0, 3)
    e.square_mul_repeat(0, 96)
    e.sync()
    n = e.get_size()
    cnt = 20000 if n <= (1 << 16) else 3000 if n <= (1 << 21) else 600
    ms = e.bench_square_mul(0, cnt)
    rows.append({"p": p, "n": n, "us": round(ms / cnt * 1000, 2), "ips": round(cnt / ms * 1000, 1)})
    print(f"TS={n:9d} p={p:10d} {ms / cnt * 1000:9.2f} us/iter {cnt / ms * 1000:10.1f} it/s", flush=True)
    e.close()
print(json.dumps({"ladder": rows}))
