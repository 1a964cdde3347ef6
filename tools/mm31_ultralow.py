"""README.md:88-98 of the reference: `prmers 2147483647 -pm1 -b1 100 -b2 5000 -pm1-ultralowmem -nogcd-stage1`, expected
factor 295257526626031.  One register of n = 5 * 2^25 digits, ~7300 fast3 squarings.  The 2^31-bit gcd is replaced by the
direct check that the known factor divides X - 1 (pass --gcd for the full gcd through libgmp; minutes of host time)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import prmers_b200 as pb
from prmers_b200 import modes

p, b1, b2, f = 2147483647, 100, 5000, 295257526626031
full = "--gcd" in sys.argv
t0 = time.time()
e = pb.Engine(p, 1, verbose=True)
r = modes.run_pm1_stage2_ultralow(e, p, b1, b2, gcd=full, progress=lambda d, n: print(f"  {d}/{n} bits, {time.time() - t0:.1f} s", flush=True))
e.sync()
t1 = time.time()
x = r["x"]
out = {"exponent": p, "b1": b1, "b2": b2, "fft_length": r["fft_length"], "bits": r["bits"], "seconds": round(t1 - t0, 1),
       "known_factor": f, "known_factor_divides_Mp": pow(2, p, f) == 1, "known_factor_divides_X_minus_1": (x - 1) % f == 0,
       "full_gcd": r["factor"] if full else "skipped"}
print(json.dumps(out))
assert out["known_factor_divides_X_minus_1"] and out["known_factor_divides_Mp"]
