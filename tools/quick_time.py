"""Quick timing probe (development aid): ms per square_mul at a few sizes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import prmers_b200 as pb

for p in [int(a) for a in sys.argv[1:]] or [136279841, 205271257, 86243, 1362763, 20000003]:
    e = pb.Engine(p, 4, verbose=True)
    e.set(0, 3)
    e.square_mul_repeat(0, 100)
    e.sync()
    cnt = int(os.environ.get("QT_COUNT", 2000 if e.get_size() >= (1 << 20) else 20000))
    ms = e.bench_square_mul(0, cnt)
    n = e.get_size()
    print(f"p={p} n={n} {ms / cnt * 1000:.2f} us/iter  {cnt / ms * 1000:.0f} it/s  48n/t = {48 * n / (ms / cnt * 1e-3) / 1e9:.0f} GB/s", flush=True)
    e.close()
