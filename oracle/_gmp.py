"""Minimal ctypes binding of the system libgmp runtime (TEST INFRASTRUCTURE ONLY).

Only what the big-int reference needs: Mersenne-reduced squaring chains that pure Python ints are
too slow for (p ~ 1e5 .. 1e8).  Falls back to Python ints when libgmp is not loadable."""
import ctypes as C
import ctypes.util


class _MPZ(C.Structure):
    _fields_ = [("alloc", C.c_int), ("size", C.c_int), ("d", C.c_void_p)]


def _load():
    name = ctypes.util.find_library("gmp") or "libgmp.so.10"
    try:
        g = C.CDLL(name)
    except OSError:
        return None
    g.__gmpz_sizeinbase.restype = C.c_size_t
    g.__gmpz_export.restype = C.c_void_p
    return g


_g = _load()


class _F:
    """Aliases without the leading double underscore (Python would mangle it inside a class)."""

    def __getattr__(self, name):
        return getattr(_g, "__gmpz_" + name)


_f = _F()


def available():
    return _g is not None


class Z:
    """A GMP integer with just enough operations for x -> x^2 * a mod 2^p - 1."""

    def __init__(self, v=0):
        self.z = _MPZ()
        _f.init(C.byref(self.z))
        if v:
            self.set(v)

    def __del__(self):
        try:
            _f.clear(C.byref(self.z))
        except Exception:
            pass

    def set(self, v):
        raw = int(v).to_bytes((int(v).bit_length() + 7) // 8 or 1, "little")
        getattr(_f, "import")(C.byref(self.z), C.c_size_t(len(raw)), C.c_int(-1), C.c_size_t(1), C.c_int(0),
                         C.c_size_t(0), raw)

    def __int__(self):
        nbytes = (_f.sizeinbase(C.byref(self.z), 2) + 7) // 8
        buf = C.create_string_buffer(nbytes + 8)
        cnt = C.c_size_t(0)
        _f.export(buf, C.byref(cnt), C.c_int(-1), C.c_size_t(1), C.c_int(0), C.c_size_t(0), C.byref(self.z))
        return int.from_bytes(buf.raw[:cnt.value], "little")


def mersenne_chain(p, x0, steps, mul_small=None, sub_small=0):
    """x <- (x^2 * a_i - sub_small) mod 2^p - 1 for i < steps; a_i = mul_small[i] (default 1).
    Returns the final value as a Python int in [0, 2^p - 1)."""
    m = (1 << p) - 1
    if _g is None:
        x = x0 % m
        for i in range(steps):
            x = x * x * (mul_small[i] if mul_small else 1) - sub_small
            x %= m
        return x
    x, t, hi, M = Z(x0 % m), Z(), Z(), Z(m)
    X, T, H, Mz = C.byref(x.z), C.byref(t.z), C.byref(hi.z), C.byref(M.z)
    pp = C.c_ulong(p)
    for i in range(steps):
        _f.mul(T, X, X)
        if mul_small is not None and mul_small[i] != 1:
            _f.mul_ui(T, T, C.c_ulong(mul_small[i]))
        while _f.sizeinbase(T, 2) > p:
            _f.tdiv_q_2exp(H, T, pp)
            _f.tdiv_r_2exp(T, T, pp)
            _f.add(T, T, H)
        if sub_small:
            if _f.cmp_ui(T, C.c_ulong(sub_small)) < 0:
                _f.add(T, T, Mz)
            _f.sub_ui(T, T, C.c_ulong(sub_small))
        _f.set(X, T)
    v = int(x)
    return 0 if v == m else v % m


def mulmod_mersenne(x, y, p):
    """x * y mod 2^p - 1 for huge operands (GMP product + fold); Python ints in and out."""
    m = (1 << p) - 1
    if _g is None or p < 200000:
        t = x * y
        while t >> p:
            t = (t & m) + (t >> p)
        return 0 if t == m else t
    a, b, t, hi = Z(x), Z(y), Z(), Z()
    T, H = C.byref(t.z), C.byref(hi.z)
    _f.mul(T, C.byref(a.z), C.byref(b.z))
    pp = C.c_ulong(p)
    while _f.sizeinbase(T, 2) > p:
        _f.tdiv_q_2exp(H, T, pp)
        _f.tdiv_r_2exp(T, T, pp)
        _f.add(T, T, H)
    v = int(t)
    return 0 if v == m else v
