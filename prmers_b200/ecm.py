"""ECM on twisted Edwards curves over the engine API (BASELINE config 5).

Host loop only: every field operation is an engine call on the register map of the reference
(RunEcmTwistedEdwards.cpp:2702-2705,3055-3090: point (X,Y,Z,T) = R3,R4,R1,R5; affine base (X2,Y2,T2) = R6,R7,R9
and its negative R47,R48,R49; constants a = R16, 2d = R17, d = R29, 2 = R8; persistent multiplicands R43 (a),
R44 (2), R45 (d), R46 (T2), R50 (-T2); scratch multiplicand R11; temporaries R20-R42; 51 registers).
Restated from src/modes/RunEcmTwistedEdwards.cpp of the reference:
  :859-886    splitmix64 / mix64 / rnd_mpz_bits (curve_seed_c = mix64(base_seed, c))
  :1399-1431  prime lists, :1455-1465 + :1500-1512 stage-1 exponent K = prod q^floor(log_q B1)
  :2436-2598  default ("no torsion") curve: Suyama sigma -> Montgomery (A, B) -> twisted Edwards (a, d, X0, Y0)
  :2717-2909  eADD_RP / eADD_RP_2 (mixed addition of +-P) and eDBL_XYTZ (doubling, general a)
  :3126-3153,3264-3283  NAF of the scalar, double-and-add from the top digit
  :2626-2652  curve-equation invariant a X^2 + Y^2 = Z^2 + d T^2
  :3372-3440  stage-1 gcd(T, N), :3542-3566 stage-2 base = affine(Q), :3569-3606,3775-3794 stage-2 chunks
  :723-803,2173-2435  the other curve families: -torsion16 (a = 1, Z/2 x Z/8 torsion) and -iv163 (family IV, a = -1)
Not restated (out of the hot path's scope): checkpoint files, Prime95 hand-off, GUI, the Montgomery-form driver.
"""
import math

from .modes import _set_int, _value

M64 = (1 << 64) - 1
REGS = 51


# ---- seeds (RunEcmTwistedEdwards.cpp:859-886) --------------------------------------------------------
def splitmix64_step(x):
    x = (x + 0x9E3779B97F4A7C15) & M64
    z = x
    z ^= z >> 30; z = (z * 0xBF58476D1CE4E5B9) & M64
    z ^= z >> 27; z = (z * 0x94D049BB133111EB) & M64
    z ^= z >> 31
    return x, z


def mix64(seed, idx):
    x = (seed ^ 0x9E3779B97F4A7C15) & M64
    x ^= ((idx + 1) * 0xBF58476D1CE4E5B9) & M64
    x ^= ((idx + 0x100) * 0x94D049BB133111EB) & M64
    return splitmix64_step(x)[1]


def rnd_mpz_bits(N, seed0, bits):
    z, s = 0, seed0 & M64
    for _ in range(0, bits, 64):
        s, w = splitmix64_step(s)
        z = (z << 64) + w
    z %= N
    return z + 3 if z <= 2 else z


# ---- scalars ----------------------------------------------------------------------------------------
def _primes_upto(n):
    if n < 2:
        return []
    sieve = bytearray([1]) * (n + 1)
    sieve[:2] = b"\0\0"
    for i in range(2, int(n ** 0.5) + 1):
        if sieve[i]:
            sieve[i * i::i] = bytearray(len(sieve[i * i::i]))
    return [i for i in range(n + 1) if sieve[i]]


def ecm_primes(b1, b2):
    """(primes <= B1, primes in (B1, B2]) -- :1399-1431."""
    ps = _primes_upto(max(b1, b2))
    return [q for q in ps if q <= b1], [q for q in ps if b1 < q <= b2]


def stage1_exponent(b1):
    """K = prod_{q <= B1} q^floor(log_q B1) (:1459-1463)."""
    k = 1
    for q in _primes_upto(b1):
        mm = q
        while mm <= b1 // q:
            mm *= q
        k *= mm
    return k


def naf(e):
    """Non-adjacent form, least significant digit first, top zeros trimmed (:3129-3141,3174)."""
    out = []
    while e:
        d = 0
        if e & 1:
            d = 1 if (e & 3) == 1 else -1
            e -= d
        out.append(d)
        e >>= 1
    while out and out[-1] == 0:
        out.pop()
    return out


def s2_chunk_bits(n):
    """compute_s2_chunk_bits (:138-144)."""
    for lim, bits in ((256, 1048576), (1024, 524288), (4096, 262144), (16384, 131072)):
        if n <= lim:
            return bits
    return 65536


def s2_chunks(primes_s2, max_bits):
    """Stage-2 scalars: products of consecutive primes, closed when they reach max_bits (:3694-3699)."""
    out, i = [], 0
    while i < len(primes_s2):
        e, j = 1, i
        while j < len(primes_s2):
            e *= primes_s2[j]; j += 1
            if e.bit_length() >= max_bits:
                break
        out.append((i, j, e))
        i = j
    return out


# ---- curve from a seed (:2436-2598) --------------------------------------------------------------------
class FactorFound(Exception):
    """A modular inversion hit a proper divisor of N (invm returning 1, :2108-2119)."""
    def __init__(self, g):
        super().__init__(str(g)); self.g = g


def _inv(a, N):
    """invm: None when a = 0 or gcd(a, N) = N, FactorFound on a proper divisor."""
    a %= N
    if a == 0:
        return None
    g = math.gcd(a, N)
    if g == 1:
        return pow(a, -1, N)
    if 1 < g < N:
        raise FactorFound(g)
    return None


_CHKSUMMOD = 4294967291


def ecm_resume_line(p, b1, A, x, program="PrMers b200", who="", when=""):
    """One GMP-ECM / Prime95 resume line for a finished stage 1 (append_ecm_stage1_resume_line, :1287-1330): the curve as
    its Montgomery coefficient A = 2 (a + d) / (a - d), the point as its Montgomery x = (Z + Y) / (Z - Y)."""
    N = (1 << p) - 1
    A, x = A % N, x % N
    chk = b1 * (A % _CHKSUMMOD) * (N % _CHKSUMMOD) * (x % _CHKSUMMOD) % _CHKSUMMOD
    return (f"METHOD=ECM; B1={b1}; N=2^{p}-1; X=0x{x:x}; A={A}; CHECKSUM={chk}; PROGRAM={program}; X0=0x0; Y0=0x0;"
            + (f" WHO={who};" if who else "") + f" TIME={when};\n")


def parse_ecm_resume_line(line):
    """The fields back, checksum verified."""
    f = dict(kv.strip().split("=", 1) for kv in line.strip().rstrip(";").split(";") if "=" in kv)
    b1, p, x, A = int(f["B1"]), int(f["N"][2:].split("-")[0]), int(f["X"], 16), int(f["A"])
    N = (1 << p) - 1
    if int(f["CHECKSUM"]) != b1 * (A % _CHKSUMMOD) * (N % _CHKSUMMOD) * (x % _CHKSUMMOD) % _CHKSUMMOD:
        raise ValueError("ECM resume line: checksum mismatch")
    return {"p": p, "b1": b1, "x": x, "A": A}


def ecm_resume_name(p, b1):
    return f"resume_p{p}_ECM_TE_B1_{b1}.save"


def curve_from_sigma(N, sigma):
    """Suyama parametrisation -> Montgomery (A + 2, B) -> twisted Edwards a = (A+2)/B, d = (A-2)/B and the point
    (X0, Y0); None when the sigma is singular for this modulus (the reference then tries the next one)."""
    sigma %= N
    if sigma == 0:
        sigma = 1
    s2 = sigma * sigma % N
    su = (s2 - 5) % N
    sv = 4 * sigma % N
    inv_sv = _inv(sv, N)
    if inv_sv is None:
        return None
    den_a = 4 * (su * su % N * su % N) * sv % N
    inv_den_a = _inv(den_a, N)
    if inv_den_a is None:
        return None
    d = (sv - su) % N
    tnum = d * d % N * d % N
    aplus2 = tnum * ((3 * su + sv) % N) % N * inv_den_a % N
    bmont = su * inv_sv % N
    inv_b = _inv(bmont, N)
    if inv_b is None:
        return None
    aE = aplus2 * inv_b % N
    dE = (aplus2 - 4) % N * inv_b % N
    sp = (s2 + 5) % N
    den_x = (su - sv) % N * ((su + sv) % N) % N * sp % N
    inv_den_x = _inv(den_x, N)
    if inv_den_x is None:
        return None
    X0 = su * su % N * sv % N * inv_den_x % N
    su3, sv3 = su * su % N * su % N, sv * sv % N * sv % N
    inv_den_y = _inv((su3 + sv3) % N, N)
    if inv_den_y is None:
        return None
    Y0 = (su3 - sv3) % N * inv_den_y % N
    if X0 == 0 or Y0 == 0:
        return None
    x2, y2 = X0 * X0 % N, Y0 * Y0 % N
    if (aE * x2 + y2 - 1 - dE * x2 % N * y2) % N != 0:
        return None
    return {"a": aE, "d": dE, "X0": X0, "Y0": Y0, "sigma": sigma}


def curve_from_seed(N, curve_seed):
    """Default curve family of the reference: up to 256 sigmas derived from the curve seed (:2436-2461)."""
    for tries in range(256):
        sigma_seed = curve_seed if tries == 0 else mix64(curve_seed, 0x5EED5EED + tries)
        cv = curve_from_sigma(N, rnd_mpz_bits(N, sigma_seed, 63))
        if cv is not None:
            cv["tries"] = tries
            return cv
    return None


# ---- the other curve families of the reference (-torsion16, -iv163) -------------------------------------------------
def _ec_mod4_mul(n, s1, t1, N):
    """ecm_local::EC_mod4::get (:723-803): n * (s1, t1) on y^2 = x^3 + 4 x - 16 ... by double-and-add from the top bit, affine
    coordinates modulo N; a non-invertible denominator ends at the point at infinity, reported as (0, 0)."""
    def inv(d):
        d %= N
        if d == 0 or math.gcd(d, N) != 1:
            return None
        return pow(d, -1, N)

    def dbl(P):
        if P is None:
            return None
        i = inv(2 * P[1])
        if i is None:
            return None
        lam = (3 * P[0] * P[0] + 4) % N * i % N
        x3 = (lam * lam - 2 * P[0]) % N
        return (x3, (lam * (P[0] - x3) - P[1]) % N)

    def add(P, Q):
        if P is None:
            return Q
        if Q is None:
            return P
        if P[0] == Q[0]:
            return None if (P[1] + Q[1]) % N == 0 else dbl(P)
        i = inv(Q[0] - P[0])
        if i is None:
            return None
        lam = (Q[1] - P[1]) % N * i % N
        x3 = (lam * lam - P[0] - Q[0]) % N
        return (x3, (lam * (P[0] - x3) - P[1]) % N)

    P0 = (s1 % N, t1 % N)
    P = P0
    for b in range(n.bit_length() - 2, -1, -1):
        P = dbl(P)
        if (n >> b) & 1:
            P = add(P, P0)
        if P is None:
            return 0, 0
    return P


def curve_torsion16(N, curve_seed):
    """-torsion16 (:2228-2370): Edwards curves with a = 1 and torsion Z/2 x Z/8 from the multiples m (4, 8) of a point of
    infinite order (Gallot / Theorem 2.5 of Bernstein-Birkner-Lange-Peters); up to 128 values of m per curve seed."""
    for tries in range(128):
        m = mix64(curve_seed, 0x544F523136 + tries) | 1
        if m < 3:
            m += 2
        s, t = _ec_mod4_mul(m, 4, 8, N)
        inv = _inv((s - 4) % N, N)
        if inv is None:
            continue
        alpha = (t + 8) % N * inv % N
        alpha2 = alpha * alpha % N
        inv = _inv((8 - alpha2) % N, N)
        if inv is None:
            continue
        r = (8 + 2 * alpha) % N * inv % N
        t1 = (2 * r - 1) % N
        t1 = t1 * t1 % N
        inv = _inv(t1 * t1 % N, N)
        if inv is None:
            continue
        dE = (8 * r * r - 8 * r + 1) % N * inv % N
        if dE in (0, 1, N - 1):
            continue
        inv = _inv((2 * s - alpha2 + 4) % N, N)
        if inv is None:
            continue
        X0 = (8 - alpha2) % N * ((2 * r * r - 1) % N) % N * inv % N
        inv = _inv((4 * r - 3) % N, N)
        if inv is None:
            continue
        Y0 = t1 * inv % N
        if X0 == 0 or Y0 == 0:
            continue
        x2, y2 = X0 * X0 % N, Y0 * Y0 % N
        if (x2 + y2 - 1 - dE * x2 % N * y2) % N != 0:
            continue
        return {"a": 1, "d": dE, "X0": X0, "Y0": Y0, "sigma": None, "tries": tries, "torsion": "16", "m": m}
    return None


def curve_family_iv(N, curve_seed):
    """-iv163 (:2180-2226,2372-2435): the Gelin-Kleinjung-Lenstra family (a = -1, 16/3 average 2-valuation): a rational
    point m (5, 8) on y^2 = x^3 - x^2 - 9 x + 9 gives t, e, and from them d = -e^4 and the starting point, computed over Q
    and only then reduced modulo N (an inversion that fails there exposes a factor)."""
    from fractions import Fraction as Q

    def padd(P, R):
        if P is None:
            return R
        if R is None:
            return P
        if P[0] == R[0] and (P[1] == -R[1] or P[1] == 0):
            return None
        if P == R:
            lam = (3 * P[0] * P[0] - 2 * P[0] - 9) / (2 * P[1])
        else:
            lam = (R[1] - P[1]) / (R[0] - P[0])
        x3 = lam * lam + 1 - P[0] - R[0]
        return (x3, -P[1] - lam * (x3 - P[0]))

    def pmul(P, k):
        R = None
        while k:
            if k & 1:
                R = padd(R, P)
            P = padd(P, P)
            k >>= 1
        return R

    def to_mod(q):
        inv = _inv(q.denominator % N, N)
        return None if inv is None else q.numerator % N * inv % N

    for tries in range(128):
        m = 1 + mix64(curve_seed, 0x163163 + tries) % 100
        Pq = pmul((Q(5), Q(8)), m)
        if Pq is None or Pq[1] == 4:
            continue
        t = (4 * Pq[0] + 4) / (Pq[1] - 4)
        t2 = t * t
        if t2 == 4:
            continue
        e = (t2 + 4 * t) / (t2 - 4)
        if e == 0:
            continue
        t4 = t2 * t2
        t6 = t4 * t2
        den_x = t4 + 6 * t * t2 + 12 * t2 + 16 * t
        den_y = t6 + 6 * t4 * t + 10 * t4 + 16 * t * t2 + 48 * t2 + 64 * t
        if den_x == 0 or den_y == 0:
            continue
        Xq = (2 * t * t2 + 2 * t2 - 8 * t - 8) / den_x
        Yq = (t6 + 6 * t4 * t + 10 * t4 - 16 * t * t2 - 48 * t2 - 32 * t - 32) / den_y
        dE, X0, Y0 = to_mod(-(e ** 4)), to_mod(Xq), to_mod(Yq)
        if dE is None or X0 is None or Y0 is None:
            continue
        aE = N - 1
        if dE in (0, 1, N - 1) or X0 == 0 or Y0 == 0:
            continue
        x2, y2 = X0 * X0 % N, Y0 * Y0 % N
        if (aE * x2 + y2 - 1 - dE * x2 % N * y2) % N != 0:
            continue
        return {"a": aE, "d": dE, "X0": X0, "Y0": Y0, "sigma": None, "tries": tries, "torsion": "family_iv_163", "m": m}
    return None


def curve_for(N, curve_seed, torsion=None):
    """The curve of one ECM run: default family (Suyama sigma), "16" (-torsion16) or "family_iv_163" (-iv163)."""
    if torsion in (None, "none"):
        return curve_from_seed(N, curve_seed)
    if torsion == "16":
        return curve_torsion16(N, curve_seed)
    if torsion == "family_iv_163":
        return curve_family_iv(N, curve_seed)
    raise ValueError(f"unknown curve family {torsion!r}")


# ---- point operations on engine registers ------------------------------------------------------------
class EdwardsOps:
    """The reference's op schedules, register for register.  `eng` needs the engine.h methods; the fused helpers
    (addsub, addsub_copy, mul_copy) fall back to their default compositions (engine.h:65-131) when absent."""

    def __init__(self, eng, notwist=False, graphs=None):
        self.e = eng
        self.ops = 0
        self.notwist = notwist
        # B200: a point operation is 17-27 engine calls of a few microseconds of launch latency each at ECM sizes, so
        # record the three schedules once and replay them (one kernel launch per point operation instead of one per op)
        if graphs is None:
            graphs = hasattr(eng, "graph_begin") and bool(getattr(eng, "is_single_cta", lambda: False)())
        self._g = None
        if graphs:
            self._g = {}
            for key, fn in (("dbl", lambda: self._dbl()), ("add+", lambda: self._add(False)), ("add-", lambda: self._add(True))):
                eng.graph_begin()
                try:
                    fn()
                finally:
                    self._g[key] = eng.graph_end()

    def dbl(self):
        self.ops += 1
        if self._g:
            self.e.graph_launch(self._g["dbl"])
        else:
            self._dbl()

    def add(self, negative=False):
        self.ops += 1
        if self._g:
            self.e.graph_launch(self._g["add-" if negative else "add+"])
        else:
            self._add(negative)

    def _dbl(self):
        (self.dbl_notwist if self.notwist else self.dbl_twisted)(3, 4, 1, 5)

    def _add(self, negative):
        (self.add_notwist if self.notwist else self.add_twisted)(negative)

    def _addsub(self, s, d, a, b):
        e = self.e
        if hasattr(e, "addsub"):
            e.addsub(s, d, a, b)
        else:
            e.copy(s, a); e.copy(d, a); e.add(s, b); e.sub_reg(d, b)

    def _addsub_copy(self, s, d, sc, dc, a, b):
        e = self.e
        if hasattr(e, "addsub_copy"):
            e.addsub_copy(s, d, sc, dc, a, b)
        else:
            self._addsub(s, d, a, b); e.copy(sc, s); e.copy(dc, d)

    def _mul_copy(self, dst, src, cp):
        e = self.e
        if hasattr(e, "mul_copy"):
            e.mul_copy(dst, src, cp)
        else:
            e.mul(dst, src); e.copy(cp, dst)

    def dbl_twisted(self, RX=3, RY=4, RZ=1, RT=5):
        """eDBL_XYTZ (:2867-2909): A = X^2, B = Y^2, C = 2 Z^2, D = a A, E = 2 T Z, G = D + B, H = D - B, F = G - C;
        X3 = E F, Y3 = G H, T3 = E H, Z3 = F G."""
        e = self.e
        e.set_multiplicand(11, RZ); e.mul(RT, 11); e.add(RT, RT)
        e.square_mul(RZ); e.add(RZ, RZ)
        e.square_mul(RX); e.square_mul(RY)
        e.mul(RX, 43)
        self._addsub(23, 25, RX, RY)
        e.copy(24, 23); e.sub_reg(24, RZ)
        e.set_multiplicand(11, 24)
        e.copy(RX, RT); e.mul(RX, 11)
        e.copy(RZ, 23); e.mul(RZ, 11)
        e.set_multiplicand(11, 25)
        e.copy(RY, 23); e.mul(RY, 11)
        e.mul(RT, 11)

    def dbl_notwist(self, RX=3, RY=4, RZ=1, RT=5):
        """eDBL_XYTZ_notwist (:2911-2945), a = 1."""
        e = self.e
        e.set_multiplicand(11, RZ); e.mul(RT, 11); e.add(RT, RT)
        e.square_mul(RZ); e.add(RZ, RZ)
        e.square_mul(RX); e.square_mul(RY)
        self._addsub_copy(23, 25, 24, 22, RX, RY)
        e.sub_reg(24, RZ)
        e.set_multiplicand(11, 24)
        e.copy(RX, RT); e.mul(RX, 11)
        e.copy(RZ, 23); e.mul(RZ, 11)
        e.set_multiplicand(11, 25)
        e.copy(RY, 23); e.mul(RY, 11)
        e.mul(RT, 11)

    def add_twisted(self, negative=False):
        """eADD_RP / eADD_RP_2 (:2717-2772,2794-2848): mixed addition of the affine base (Z2 = 1) or of its negative.
        E = X1 Y2 + Y1 X2, H = Y1 Y2 - a X1 X2, C = d T1 T2, D = Z1; X3 = E (D - C), Y3 = H (D + C), T3 = E H,
        Z3 = (D - C)(D + C)."""
        e = self.e
        X2, Y2, MT2 = (47, 48, 50) if negative else (6, 7, 46)
        self._addsub(34, 35, 4, 3)
        self._addsub(36, 37, Y2, X2)
        e.copy(30, 3); e.set_multiplicand(11, X2); self._mul_copy(30, 11, 39)
        e.copy(31, 4); e.set_multiplicand(11, Y2); e.mul(31, 11)
        e.copy(32, 5); e.mul(32, MT2); e.mul(32, 45)
        self._addsub(42, 41, 1, 32)
        e.copy(38, 34); e.set_multiplicand(11, 36); e.mul(38, 11)
        e.sub_reg(38, 30); e.sub_reg(38, 31)
        e.mul(39, 43)
        e.copy(40, 31); e.sub_reg(40, 39)
        e.copy(3, 38); e.set_multiplicand(11, 41); e.mul(3, 11)
        e.copy(1, 42); e.mul(1, 11)
        e.set_multiplicand(11, 40)
        e.copy(4, 42); e.mul(4, 11)
        e.copy(5, 38); e.mul(5, 11)

    def add_notwist(self, negative=False):
        """eADD_RP_notwist / _2 (:2957-3067), a = 1."""
        e = self.e
        X2, Y2, MT2 = (47, 48, 50) if negative else (6, 7, 46)
        self._addsub(34, 35, 4, 3)
        self._addsub(36, 37, Y2, X2)
        e.copy(30, 3); e.set_multiplicand(11, X2); e.mul(30, 11)
        e.copy(31, 4); e.set_multiplicand(11, Y2); e.mul(31, 11)
        self._addsub(20, 40, 31, 30)
        e.copy(32, 5); e.mul(32, MT2); e.mul(32, 45)
        self._addsub(42, 41, 1, 32)
        e.copy(38, 34); e.set_multiplicand(11, 36); e.mul(38, 11)
        e.sub_reg(38, 20)
        e.set_multiplicand(11, 40)
        e.copy(4, 42); e.mul(4, 11)
        e.copy(5, 38); e.mul(5, 11)
        e.set_multiplicand(11, 41)
        e.copy(3, 38); e.mul(3, 11)
        e.copy(1, 42); e.mul(1, 11)


def _load_base(eng, p, N, X, Y):
    """Base point +-P in R6,R7,R9 / R47,R48,R49 and the persistent multiplicands (:3180-3201, :1851-1864)."""
    T = X * Y % N
    for r, v in ((6, X), (7, Y), (9, T), (47, (N - X) % N), (48, Y), (49, (N - T) % N)):
        _set_int(eng, r, v, p)
    eng.set_multiplicand(43, 16); eng.set_multiplicand(44, 8); eng.set_multiplicand(45, 29)
    eng.set_multiplicand(46, 9); eng.set_multiplicand(50, 49)


def _scalar_mul(eng, ops, digits, on_step=None, start_step=0):
    """Q <- k P over the NAF digits (top digit first); the point registers already hold +-P for the top digit.
    `start_step`: steps already done by a run whose checkpoint has been restored."""
    for step, i in enumerate(range(len(digits) - 2, -1, -1)):
        if step < start_step:
            continue
        ops.dbl()
        d = digits[i]
        if d:
            ops.add(negative=d < 0)
        if on_step:
            on_step(step + 1)


def check_invariant(eng, p, aE, dE):
    """a X^2 + Y^2 == Z^2 + d T^2 on the current point (:2626-2652)."""
    N = (1 << p) - 1
    X, Y, Z, T = (_value(eng, r, p) for r in (3, 4, 1, 5))
    return (aE * X * X + Y * Y - Z * Z - dE * T * T) % N == 0


def run_curve(eng, p, b1, b2, curve_seed, curve=None, K=None, primes_s2=None, error_iter=0, log=None, graphs=None, torsion=None,
              checkpoint_path=None, backup_every=0, stop_at=None, checkpoint2_path=None, stop_at2=None, s2_bits=None, resume_path=None):
    """One curve of RunEcmTwistedEdwards.cpp:1737-3859 on a 51-register engine.  Returns a dict with `status` in
    {"found", "NF", "singular", "nocurve", "invariant"}, `factor`, `stage` (0 = curve generation, 1, 2) and the final
    extended coordinates (for parity checks against a big-integer restatement of the same formulas).
    Stage-1 checkpoints (:1866-1937, :3090-3243): with `checkpoint_path` the scalar multiplication resumes from that file
    when it holds this exponent, bound, curve seed and family, saves every `backup_every` steps and, after `stop_at`
    steps (the reference's SIGINT path), saves and returns status "interrupted".
    Stage-2 checkpoints (:1940-2078, :3092, :3640-3664, :3797-3822): with `checkpoint2_path` a stage 2 that a file of this
    exponent, bounds, curve seed and family describes is resumed instead of running stage 1 again -- at a chunk boundary,
    or inside a chunk at the saved step with the saved affine base; the state is saved after every chunk, every
    `backup_every` steps inside a chunk and at `stop_at2` = (chunk number, steps) before returning "interrupted"; a
    finished stage 2 removes the file.  `s2_bits` overrides the chunk size (tests).
    `resume_path`: a stage 1 that ends without a factor appends its GMP-ECM resume line there (:1287-1330, :3396-3426);
    the two values are also returned as `resume_x` and `resume_A`."""
    say = log or (lambda s: None)
    N = (1 << p) - 1
    res = {"curve_seed": curve_seed, "status": "NF", "factor": None, "stage": None, "n": eng.get_size()}
    try:
        cv = curve if curve is not None else curve_for(N, curve_seed, torsion)
    except FactorFound as f:
        res.update(status="found", factor=f.g, stage=0)
        return res
    if cv is None:
        res["status"] = "nocurve"
        return res
    aE, dE, X0, Y0 = cv["a"], cv["d"], cv["X0"], cv["Y0"]
    res["sigma"] = cv.get("sigma")
    say(f"[ECM] curve_seed={curve_seed} | twisted_edwards | torsion=none")

    # constants and the starting point (:3069-3090)
    _set_int(eng, 16, aE, p); _set_int(eng, 17, 2 * dE % N, p); _set_int(eng, 29, dE, p)
    eng.set(0, 0); eng.set(1, 1); eng.set(8, 2)
    if K is None:
        K = stage1_exponent(b1)
    digits = naf(K)
    _load_base(eng, p, N, X0, Y0)
    neg = bool(digits) and digits[-1] < 0
    T0 = X0 * Y0 % N
    for r, v in ((3, (N - X0) % N if neg else X0), (4, Y0), (5, (N - T0) % N if neg else T0)):
        _set_int(eng, r, v, p)

    # -torsion16 curves have a = 1 and run the "notwist" schedules (:3255-3272); every other family the general-a ones
    ops = EdwardsOps(eng, notwist=(torsion == "16"), graphs=graphs)
    def inject(step):
        # -erroriter (:3285-3288): corrupt Z after that many steps; the invariant must catch it
        if step == error_iter:
            eng.sub(1, 2)
            say(f"Injected error at iteration {step}")
    start_step = 0
    if checkpoint_path or checkpoint2_path:
        from . import checkpoint as ckpt
    if b2 > b1 and primes_s2 is None:
        primes_s2 = ecm_primes(b1, b2)[1]
    resume2 = None
    if checkpoint2_path and b2 > b1:
        resume2 = ckpt.load_ecm2(checkpoint2_path, eng, p, len(primes_s2), b1, b2, curve_seed, torsion)
        if resume2 is not None:
            say(f"Resuming stage 2 at prime index {resume2['idx']}" + (f", step {resume2['chunk'][3]} of its chunk." if resume2["chunk"] else "."))
    if checkpoint_path and resume2 is None:
        got = ckpt.load_ecm(checkpoint_path, eng, p, K.bit_length(), b1, curve_seed, torsion)
        if got is not None:
            start_step = got
            say(f"Resuming from a checkpoint at step {start_step}.")

    class _Interrupted(Exception):
        pass

    def on_step(step):
        if error_iter:
            inject(step)
        if checkpoint_path and ((stop_at is not None and step == stop_at) or (backup_every and step % backup_every == 0)):
            ckpt.save_ecm(checkpoint_path, eng, p, step, K.bit_length(), b1, curve_seed, torsion)
            if stop_at is not None and step == stop_at:
                raise _Interrupted()
    try:
        if resume2 is None:
            _scalar_mul(eng, ops, digits, on_step=on_step if (error_iter or checkpoint_path) else None, start_step=start_step)
    except _Interrupted:
        say(f"Interrupted by user, state saved at step {stop_at}")
        res.update(status="interrupted", stage=1, step=stop_at)
        return res
    if resume2 is None:
        res["b1"] = b1
        early = _after_stage1(eng, p, N, aE, dE, ops, res, say, resume_path)
        if early is not None:
            return early

    if b2 > b1:
        chunks = s2_chunks(primes_s2, s2_bits or s2_chunk_bits(eng.get_size()))
        for ci, (i0, i1, echunk) in enumerate(chunks):
            if resume2 is not None and i0 < resume2["idx"]:
                continue                                         # chunks a saved state is already past
            d2 = naf(echunk)
            in_chunk = resume2 is not None and resume2["chunk"] is not None and resume2["chunk"][0] == i0
            if in_chunk:
                # the registers (point, base, multiplicands) are the saved ones; the saved base is what they were made from
                base, step0 = resume2["base"], resume2["chunk"][3]
            else:
                # setup_te_stage2_base (:3542-3566): the current point, made affine, becomes the base
                Xv, Yv, Zv = (_value(eng, r, p) for r in (3, 4, 1))
                try:
                    iz = _inv(Zv, N)
                except FactorFound as f:
                    res.update(status="found", factor=f.g, stage=2)
                    return res
                if iz is None:
                    res.update(status="singular", stage=2)
                    return res
                Xa, Ya = Xv * iz % N, Yv * iz % N
                Ta = Xa * Ya % N
                base, step0 = [Xa, Ya, Ta, (N - Xa) % N, Ya, (N - Ta) % N], 0
                _load_base(eng, p, N, Xa, Ya)
                if d2:
                    eng.set(1, 1)
                    for dst, src in (((3, 47), (4, 48), (5, 49)) if d2[-1] < 0 else ((3, 6), (4, 7), (5, 9))):
                        eng.copy(dst, src)
            resume2 = None

            def on_step2(step, ci=ci, i0=i0, i1=i1, nb=len(d2), base=base):
                stop = stop_at2 is not None and stop_at2 == (ci, step)
                if stop or (backup_every and step % backup_every == 0):
                    ckpt.save_ecm2(checkpoint2_path, eng, p, i0, len(primes_s2), b1, b2, curve_seed, torsion, chunk=(i0, i1, nb, step), base=base)
                    if stop:
                        raise _Interrupted()
            if d2:
                try:
                    _scalar_mul(eng, ops, d2, on_step=on_step2 if checkpoint2_path else None, start_step=step0)
                except _Interrupted:
                    say(f"Interrupted by user in stage 2, state saved at step {stop_at2[1]} of chunk {ci}")
                    res.update(status="interrupted", stage=2, step=stop_at2[1], chunk=ci)
                    return res
            T = _value(eng, 5, p)
            g = math.gcd(T, N)
            res["stage2_T"] = T
            res["point_ops"] = ops.ops
            say(f"[ECM] Stage2 chunk end | next_prime_idx={i1} | gcd(T,N)={g}")
            if g == N:
                res.update(status="singular", stage=2)
            elif g > 1:
                res.update(status="found", factor=g, stage=2)
            if res["status"] != "NF" or ci == len(chunks) - 1:
                if checkpoint2_path:
                    ckpt.remove_ecm2(checkpoint2_path)
                return res
            if checkpoint2_path:
                ckpt.save_ecm2(checkpoint2_path, eng, p, i1, len(primes_s2), b1, b2, curve_seed, torsion)
                if stop_at2 is not None and stop_at2 == (ci, None):
                    say(f"Interrupted by user after stage-2 chunk {ci}")
                    res.update(status="interrupted", stage=2, chunk=ci)
                    return res
    return res


def _after_stage1(eng, p, N, aE, dE, ops, res, say, resume_path=None):
    """Invariant, gcd on T and the two inversions that end stage 1 (:3372-3413): the finished result, or None to go on."""
    if not check_invariant(eng, p, aE, dE):
        say("[ECM] invariant FAIL")
        res.update(status="invariant", stage=1, point_ops=ops.ops)
        return res
    say("[ECM] check invariant OK")

    # stage-1 gcd on T (:3372-3374)
    T = _value(eng, 5, p)
    g = math.gcd(T, N)
    res["stage1_T"] = T
    res["point_ops"] = ops.ops
    if g == N:
        res.update(status="singular", stage=1)
        return res
    if 1 < g < N:
        res.update(status="found", factor=g, stage=1)
        return res
    # g == 1: the Montgomery resume values need 1/(Z - Y) and 1/(a - d); either inversion may split N (:3396-3413)
    Zv, Yv = _value(eng, 1, p), _value(eng, 4, p)
    try:
        iu = _inv((Zv - Yv) % N, N)
        if iu is not None:
            ia = _inv((aE - dE) % N, N)
            if ia is not None:
                res["resume_x"] = (Zv + Yv) * iu % N          # Montgomery x of the stage-1 point
                res["resume_A"] = 2 * (aE + dE) * ia % N      # Montgomery coefficient of the curve
                if resume_path:
                    with open(resume_path, "a") as f:
                        f.write(ecm_resume_line(p, res.get("b1", 0), res["resume_A"], res["resume_x"]))
    except FactorFound as f:
        res.update(status="found", factor=f.g, stage=1)
        return res
    return None


def run_ecm(engine_factory, p, b1, b2=0, curves=1, seed=1, known_factors=(), continue_after_factor=False,
            curve_indices=None, log=None, graphs=None, torsion=None, checkpoint_dir=None, backup_every=0, stop=None, s2_bits=None,
            resume_path=None):
    """The curve loop (:1737-1830, 3415-3455): curve c uses curve_seed = mix64(seed, c) on its own engine
    (`engine_factory(p, 51)`), stops at the first new factor unless told to continue.  `curve_indices` restricts
    the loop to a subset (multi-GPU: curve c -> rank c mod world).
    With `checkpoint_dir` every curve keeps the reference's two files there (`ecm_te_m_<p>_c<c>.ckpt`,
    `ecm2_te_m_<p>_c<c>.ckpt`, :1826-1830): a curve whose files describe a stage 1 or stage 2 under way resumes from them,
    and `stop` = (curve, 1, step) or (curve, 2, (chunk, step)) is the interruption of the reference's SIGINT path: the
    state is saved, the loop ends with status "interrupted" and the same call made again carries on from the files.
    `resume_path`: every curve whose stage 1 ends without a factor appends its GMP-ECM resume line to that file
    (`resume_p<p>_ECM_TE_B1_<B1>.save` in the reference, :1269-1272)."""
    say = log or (lambda s: None)
    K = stage1_exponent(b1)
    primes_s2 = ecm_primes(b1, b2)[1] if b2 > b1 else []
    known = {int(f) for f in known_factors}
    out = {"exponent": p, "worktype": "ECM", "b1": b1, "b2": b2, "seed": seed, "status": "NF", "factor": None,
           "curves": [], "k_bits": K.bit_length()}
    for c in (range(curves) if curve_indices is None else curve_indices):
        eng = engine_factory(p, REGS)
        kw = {}
        if checkpoint_dir:
            import os
            from . import checkpoint as ckpt
            kw = {"checkpoint_path": os.path.join(checkpoint_dir, ckpt.ecm_checkpoint_name(p, c)),
                  "checkpoint2_path": os.path.join(checkpoint_dir, ckpt.ecm2_checkpoint_name(p, c)), "backup_every": backup_every}
            if stop is not None and stop[0] == c:
                kw["stop_at" if stop[1] == 1 else "stop_at2"] = stop[2]
        try:
            r = run_curve(eng, p, b1, b2, mix64(seed, c), K=K, primes_s2=primes_s2, log=log, graphs=graphs, torsion=torsion, s2_bits=s2_bits,
                          resume_path=resume_path, **kw)
        finally:
            if hasattr(eng, "close"):
                eng.close()
        r["curve"] = c
        out["curves"].append(r)
        out["fft_length"] = r["n"]
        if r["status"] == "interrupted":
            out.update(status="interrupted", interrupted_curve=c)
            break
        if checkpoint_dir and os.path.exists(kw["checkpoint_path"]):
            os.remove(kw["checkpoint_path"])                  # a finished curve leaves no stage-1 file behind (:3440-3446)
        if r["status"] == "found":
            if r["factor"] in known:
                say(f"[ECM] Curve {c + 1}/{curves} | known factor={r['factor']}")
                continue
            say(f"[ECM] Curve {c + 1}/{curves} | factor={r['factor']}")
            if out["factor"] is None:
                out.update(status="found", factor=r["factor"], curves_tested_for_found=c + 1)
            if not continue_after_factor:
                break
            known.add(r["factor"])
    return out


def run_ecm_batched(batch_engine_factory, p, b1, b2=0, curves=1, seed=1, known_factors=(), curve_indices=None, log=None, graphs=None):
    """Config 5 the B200 way: ALL curves of the call at once on ONE batched engine (`batch_engine_factory(p, 51, B)`,
    one register slab and one CTA per curve in every launch).  The stage-1 scalar K and the stage-2 chunk scalars are
    the same for every curve, so the curves run the very same sequence of point operations in lockstep -- a point
    operation is still one launch, now for B curves; what differs per curve (the curve constants, the base point, the
    gcds and inversions between the stages) is host work on one slab's digits.  Every curve's residues, status, stage
    and factor are those run_curve gives it alone (a curve that has finished early keeps being carried along and is
    ignored).  The reference runs the curves one after the other, one engine each (RunEcmTwistedEdwards.cpp:1736-1830)."""
    say = log or (lambda s: None)
    N = (1 << p) - 1
    idx = list(range(curves) if curve_indices is None else curve_indices)
    K = stage1_exponent(b1)
    primes_s2 = ecm_primes(b1, b2)[1] if b2 > b1 else []
    out = {"exponent": p, "worktype": "ECM", "b1": b1, "b2": b2, "seed": seed, "status": "NF", "factor": None,
           "curves": [], "k_bits": K.bit_length(), "batch": len(idx)}
    if not idx:
        return out
    eng = batch_engine_factory(p, REGS, len(idx))
    views = [eng.slab(s) if len(idx) > 1 else eng for s in range(len(idx))]
    res, cvs = [], []
    for s, c in enumerate(idx):
        r = {"curve": c, "curve_seed": mix64(seed, c), "status": "NF", "factor": None, "stage": None, "n": eng.get_size()}
        try:
            cv = curve_from_seed(N, r["curve_seed"])
        except FactorFound as f:
            r.update(status="found", factor=f.g, stage=0); cv = None
        else:
            if cv is None:
                r["status"] = "nocurve"
            else:
                r["sigma"] = cv.get("sigma")
        res.append(r); cvs.append(cv)
    alive = [cv is not None for cv in cvs]

    def load_base(s, X, Y):
        T = X * Y % N
        for rg, v in ((6, X), (7, Y), (9, T), (47, (N - X) % N), (48, Y), (49, (N - T) % N)):
            _set_int(views[s], rg, v, p)

    def base_multiplicands():
        eng.set_multiplicand(43, 16); eng.set_multiplicand(44, 8); eng.set_multiplicand(45, 29)
        eng.set_multiplicand(46, 9); eng.set_multiplicand(50, 49)

    def finish(s, **kw):
        res[s].update(**kw); alive[s] = False

    # constants and the starting points (:3069-3090); the small constants are the same in every slab
    eng.set(0, 0); eng.set(1, 1); eng.set(8, 2)
    digits = naf(K)
    neg = bool(digits) and digits[-1] < 0
    for s, cv in enumerate(cvs):
        if cv is None:
            continue
        aE, dE, X0, Y0 = cv["a"], cv["d"], cv["X0"], cv["Y0"]
        _set_int(views[s], 16, aE, p); _set_int(views[s], 17, 2 * dE % N, p); _set_int(views[s], 29, dE, p)
        load_base(s, X0, Y0)
        T0 = X0 * Y0 % N
        for rg, v in ((3, (N - X0) % N if neg else X0), (4, Y0), (5, (N - T0) % N if neg else T0)):
            _set_int(views[s], rg, v, p)
    base_multiplicands()
    ops = EdwardsOps(eng, graphs=graphs)
    _scalar_mul(eng, ops, digits)
    for s, cv in enumerate(cvs):
        if not alive[s]:
            continue
        v = views[s]
        if not check_invariant(v, p, cv["a"], cv["d"]):
            say(f"[ECM] curve {idx[s]}: invariant FAIL")
            finish(s, status="invariant", stage=1, point_ops=ops.ops)
            continue
        T = _value(v, 5, p)
        g = math.gcd(T, N)
        res[s]["stage1_T"] = T; res[s]["point_ops"] = ops.ops
        if g == N:
            finish(s, status="singular", stage=1); continue
        if 1 < g < N:
            finish(s, status="found", factor=g, stage=1); continue
        Zv, Yv = _value(v, 1, p), _value(v, 4, p)
        try:
            if _inv((Zv - Yv) % N, N) is not None:
                _inv((cv["a"] - cv["d"]) % N, N)
        except FactorFound as f:
            finish(s, status="found", factor=f.g, stage=1)

    if b2 > b1 and any(alive):
        for (i0, i1, echunk) in s2_chunks(primes_s2, s2_chunk_bits(eng.get_size())):
            if not any(alive):
                break
            d2 = naf(echunk)
            for s in range(len(idx)):
                if not alive[s]:
                    continue
                v = views[s]
                Xv, Yv, Zv = (_value(v, rg, p) for rg in (3, 4, 1))
                try:
                    iz = _inv(Zv, N)
                except FactorFound as f:
                    finish(s, status="found", factor=f.g, stage=2); continue
                if iz is None:
                    finish(s, status="singular", stage=2); continue
                Xa, Ya = Xv * iz % N, Yv * iz % N
                load_base(s, Xa, Ya)
            base_multiplicands()
            if d2:
                eng.set(1, 1)
                for dst, src in (((3, 47), (4, 48), (5, 49)) if d2[-1] < 0 else ((3, 6), (4, 7), (5, 9))):
                    eng.copy(dst, src)
                _scalar_mul(eng, ops, d2)
            for s in range(len(idx)):
                if not alive[s]:
                    continue
                T = _value(views[s], 5, p)
                g = math.gcd(T, N)
                res[s]["stage2_T"] = T; res[s]["point_ops"] = ops.ops
                if g == N:
                    finish(s, status="singular", stage=2)
                elif g > 1:
                    finish(s, status="found", factor=g, stage=2)
    if hasattr(eng, "launch_count"):
        out["gpu_launches"] = int(eng.launch_count())
    if hasattr(eng, "close"):
        eng.close()
    known = {int(f) for f in known_factors}
    out["curves"] = res
    for r in res:
        if r["status"] == "found" and r["factor"] not in known and out["status"] != "found":
            out.update(status="found", factor=r["factor"], curves_tested_for_found=r["curve"] + 1)
    return out


def run_ecm_sharded(engine_factory, p, b1, b2=0, curves=1, seed=1, known_factors=(), log=None, graphs=None, batch_engine_factory=None):
    """Config 5 across the GPUs of one box: curve c runs on rank c mod world with its own engine, every
    rank runs ALL of its curves (no cross-GPU stop flag: a curve at these sizes is seconds), the per-curve results are
    gathered and merged in curve order, so the answer -- first new factor by curve index -- is the one a single
    process running the curves in sequence reports."""
    from . import multi
    rank, world = multi.rank_world()
    mine = multi.assign(list(range(curves)), rank, world)
    if batch_engine_factory is not None:       # this rank's curves in lockstep on one batched engine (one CTA per curve)
        local = run_ecm_batched(batch_engine_factory, p, b1, b2, curves=curves, seed=seed, known_factors=known_factors,
                                curve_indices=mine, log=log, graphs=graphs)
    else:
        local = run_ecm(engine_factory, p, b1, b2, curves=curves, seed=seed, known_factors=known_factors,
                        continue_after_factor=True, curve_indices=mine, log=log, graphs=graphs)
    for c in local["curves"]:
        c["rank"] = rank
    allc = sorted(multi.gather(local["curves"]), key=lambda c: c["curve"])
    out = {k: local[k] for k in ("exponent", "worktype", "b1", "b2", "seed", "k_bits")}
    out.update(status="NF", factor=None, curves=allc, world=world)
    if "gpu_launches" in local:
        out["gpu_launches"] = local["gpu_launches"]              # this rank's launches
    known = {int(f) for f in known_factors}
    for c in allc:
        if c["status"] == "found" and c["factor"] not in known:
            out.update(status="found", factor=c["factor"], curves_tested_for_found=c["curve"] + 1)
            break
    return out


# ---- big-integer restatement of the same formulas (used by tests as the arithmetic reference) --------
def reference_scalar_mul(N, aE, dE, X2, Y2, digits):
    """(X, Y, Z, T) after the same double-and-add in Python integers: the engine must reproduce these residues."""
    p = N.bit_length()

    def M(x):                                         # x mod 2^p - 1 by folding (Python's % is quadratic)
        neg = x < 0
        if neg:
            x = -x
        while x >> p:
            x = (x & N) + (x >> p)
        if x == N:
            x = 0
        return (N - x) % N if neg and x else x

    T2 = M(X2 * Y2)
    sgn = -1 if digits[-1] < 0 else 1
    X, Y, Z, T = M(sgn * X2), Y2, 1, M(sgn * T2)
    for i in range(len(digits) - 2, -1, -1):
        A, B, C = M(X * X), M(Y * Y), M(2 * Z * Z)
        D, E = M(aE * A), M(2 * T * Z)
        G, H = M(D + B), M(D - B)
        F = M(G - C)
        X, Y, T, Z = M(E * F), M(G * H), M(E * H), M(F * G)
        d = digits[i]
        if d:
            x2, t2 = M(d * X2), M(d * T2)
            E = M(X * Y2 + Y * x2)
            H = M(Y * Y2 - M(aE * X) * x2)
            C = M(M(dE * T) * t2)
            X, Y, T, Z = M(E * (Z - C)), M(H * (Z + C)), M(E * H), M((Z - C) * (Z + C))
    return X, Y, Z, T
