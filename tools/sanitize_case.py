"""Small workload for compute-sanitizer (memcheck / racecheck): every kernel family once or twice."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import prmers_b200 as pb

def run(p, fs=0, r2=0):
    e = pb.Engine(p, 6, force_S=fs, force_R2=r2)
    e.set(0, 3); e.set(1, 5)
    for _ in range(3): e.square_mul(0)
    e.square_mul(0, 3)
    e.set_multiplicand(2, 1); e.mul(0, 2); e.add(0, 1); e.sub_reg(0, 1); e.sub(0, 2); e.copy(3, 0)
    v = e.get_mpz(0)
    e.close()
    return v

m = lambda p: (1 << p) - 1
def expect(p):
    x = 3
    for _ in range(3): x = x * x % m(p)
    x = x * x * 3 % m(p); x = x * 5 % m(p); x = (x + 5 - 5 - 2) % m(p)
    return x

# single CTA; forced multi-pass splits (generic kernels, second strided level); the planner's radix-4 lists; rows of 8192 and
# 10240 pairs (k_mid_fast shapes 6, 7); the streaming second level (k_level1_r4); the specialised R = 1024 strided kernels
CASES = [(127, 0, 0), (9941, 0, 0), (9941, 16, 0), (86243, 64, 0), (86243, 8, 16), (216091, 64, 0), (1362763, 0, 0),
         (86243, 0, 0), (1362763, 8192, 0), (788023, 10240, 0), (5000011, 256, 4), (40000003, 0, 0)]
for p, fs, r2 in CASES:
    assert run(p, fs, r2) == expect(p), (p, fs, r2)
print("sanitize case ok")
