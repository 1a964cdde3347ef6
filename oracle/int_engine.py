"""Register engine on plain Python integers modulo 2^p - 1 (TEST INFRASTRUCTURE ONLY).

Same method names and semantics as include/marin/engine.h:37-131 of the reference (multiplicand registers are
ordinary slots, mul never modifies its multiplicand, add/addsub aliasing works), no transform at all: what the
mode drivers compute when every engine op is exact arithmetic.  Used to pin the ECM / P-1 drivers running on the
CUDA engine (same seeds -> same residues in every register, same factors) and to generate tests/golden/.
"""


class IntEngine:
    def __init__(self, p, regs):
        self.p, self.N = p, (1 << p) - 1
        self.r = [0] * regs
        self.reg_count = regs
        self.calls = 0

    def get_size(self):
        from . import marin_oracle as mo
        return mo.transform_size(self.p)

    def close(self): pass
    def sync(self): pass

    def set(self, d, a): self.r[d] = a % self.N; self.calls += 1
    def copy(self, d, s): self.r[d] = self.r[s]; self.calls += 1
    def square_mul(self, r, a=1): self.r[r] = self.r[r] * self.r[r] * a % self.N; self.calls += 1
    def set_multiplicand(self, d, s): self.r[d] = self.r[s]; self.calls += 1
    def mul(self, d, s, a=1): self.r[d] = self.r[d] * self.r[s] * a % self.N; self.calls += 1
    def add_multiplicand(self, d, a, b): self.r[d] = (self.r[a] + self.r[b]) % self.N; self.calls += 1
    def max_multiplier(self): return 1 << 16
    def sub(self, r, a): self.r[r] = (self.r[r] - a) % self.N; self.calls += 1
    def add(self, d, s): self.r[d] = (self.r[d] + self.r[s]) % self.N; self.calls += 1
    def sub_reg(self, d, s): self.r[d] = (self.r[d] - self.r[s]) % self.N; self.calls += 1

    def addsub(self, s, d, a, b):
        x, y = self.r[a], self.r[b]
        self.r[s], self.r[d] = (x + y) % self.N, (x - y) % self.N; self.calls += 1

    def addsub_copy(self, s, d, sc, dc, a, b):
        self.addsub(s, d, a, b); self.r[sc], self.r[dc] = self.r[s], self.r[d]

    def mul_copy(self, d, s, c, a=1): self.mul(d, s, a); self.r[c] = self.r[d]
    def square_mul_copy(self, r, c, a=1): self.square_mul(r, a); self.r[c] = self.r[r]
    def mul_add(self, d, ms, as_, a=1): self.mul(d, ms, a); self.add(d, as_)
    def is_equal(self, a, b): return self.r[a] % self.N == self.r[b] % self.N
    def set_mpz(self, d, z): self.r[d] = z % self.N
    def get_mpz(self, s): return self.r[s] % self.N

    def get(self, s):
        """canonical encoded digits, like engine::get"""
        import numpy as np
        from prmers_b200.engine import int_to_digits, _widths
        return np.asarray(int_to_digits(self.r[s] % self.N, _widths(self.get_size(), self.p)), dtype=np.uint64)
