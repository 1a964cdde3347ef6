"""Host-side big-integer glue of the mode drivers: modular inverse and gcd on numbers of 10^8 bits.

The reference links GMP for exactly this (mode GCD / inverse glue: `gcd_with_dots`, AlgoUtils.hpp:338-370; V1 = H + 1/H
of the V-trace stage 2, RunPM1.cpp:3480-3520) and for nothing on the hot path; so does this module, through ctypes on
the system's libgmp runtime (no headers needed).  Python's own int gcd and pow(x, -1, m) are quadratic: fine below
~10^6 bits, hours at p = 205271257.  Without libgmp the Python versions are used whatever the size.
"""
import ctypes as C
import ctypes.util
import math


class _MPZ(C.Structure):
    _fields_ = [("alloc", C.c_int), ("size", C.c_int), ("d", C.c_void_p)]


def _load():
    for name in (ctypes.util.find_library("gmp"), "libgmp.so.10", "libgmp.so"):
        if not name:
            continue
        try:
            g = C.CDLL(name)
            g.__gmpz_sizeinbase.restype = C.c_size_t
            return g
        except OSError:
            continue
    return None


_g = _load()


def _f(name):
    """libgmp's mpz entry points are macros for __gmpz_<name>; looked up by string because Python mangles a leading
    double underscore inside a class body."""
    return getattr(_g, "__gmpz_" + name)


GMP_THRESHOLD_BITS = 200000          # below this Python's integers are fast enough


def have_gmp():
    return _g is not None


class _Z:
    def __init__(self, v=0):
        self.z = _MPZ()
        _f("init")(C.byref(self.z))
        if v:
            raw = int(v).to_bytes((int(v).bit_length() + 7) // 8, "little")
            _f("import")(C.byref(self.z), C.c_size_t(len(raw)), C.c_int(-1), C.c_size_t(1), C.c_int(0), C.c_size_t(0), raw)

    def __del__(self):
        try:
            _f("clear")(C.byref(self.z))
        except Exception:
            pass

    def __int__(self):
        nbytes = (_f("sizeinbase")(C.byref(self.z), 2) + 7) // 8
        buf = C.create_string_buffer(nbytes + 8)
        cnt = C.c_size_t(0)
        _f("export")(buf, C.byref(cnt), C.c_int(-1), C.c_size_t(1), C.c_int(0), C.c_size_t(0), C.byref(self.z))
        return int.from_bytes(buf.raw[:cnt.value], "little")


def gcd(a, b):
    """gcd(a, b) >= 0."""
    if _g is None or max(a.bit_length(), b.bit_length()) < GMP_THRESHOLD_BITS:
        return math.gcd(a, b)
    x, y, r = _Z(abs(a)), _Z(abs(b)), _Z()
    _f("gcd")(C.byref(r.z), C.byref(x.z), C.byref(y.z))
    return int(r)


def invert_mod(a, m):
    """a^-1 mod m; raises ValueError when a is not invertible (then gcd(a, m) is a factor of m)."""
    if _g is None or m.bit_length() < GMP_THRESHOLD_BITS:
        return pow(a, -1, m)
    x, y, r = _Z(a % m), _Z(m), _Z()
    if _f("invert")(C.byref(r.z), C.byref(x.z), C.byref(y.z)) == 0:
        raise ValueError("base is not invertible for the given modulus")
    return int(r)


def mulmod_mersenne(x, y, p):
    """x * y mod 2^p - 1 (product, then folding: no division)."""
    m = (1 << p) - 1
    if _g is None or p < GMP_THRESHOLD_BITS:
        t = x * y
    else:
        a, b, r = _Z(x), _Z(y), _Z()
        _f("mul")(C.byref(r.z), C.byref(a.z), C.byref(b.z))
        t = int(r)
    while t >> p:
        t = (t & m) + (t >> p)
    return 0 if t == m else t
