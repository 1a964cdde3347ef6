"""ctypes front-end of oracle/marin_oracle.c (TEST INFRASTRUCTURE ONLY).

`OracleEngine` mirrors the register API of the reference (`include/marin/engine.h:16-303`) on the
CPU restatement, so parity tests drive the oracle and the CUDA engine with the same calls.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmarin_oracle.so")
_REF_PATH = os.path.join(_HERE, "_ref", "libmarin_refhdr.so")

MOD_P = 0xFFFFFFFF00000001


def build(force=False):
    """Compile the C oracle (and oracle/_ref when /root/reference is present)."""
    src = os.path.join(_HERE, "marin_oracle.c")
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
    ecm_ref = os.path.join(_HERE, "_ref", "libecm_ref.so")
    if force or stale or (os.path.isdir("/root/reference/include/marin") and not (os.path.exists(_REF_PATH) and os.path.exists(ecm_ref))):
        subprocess.run(["make", "-C", _HERE, "all"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        u64, u32, sz, vp = C.c_uint64, C.c_uint32, C.c_size_t, C.c_void_p
        for name in ("mo_mod_add", "mo_mod_sub", "mo_mod_mul", "mo_mod_pow"):
            getattr(L, name).restype = u64
            getattr(L, name).argtypes = [u64, u64]
        L.mo_mod_inv.restype = u64; L.mo_mod_inv.argtypes = [u64]
        L.mo_root_nth.restype = u64; L.mo_root_nth.argtypes = [u64]
        L.mo_adc.restype = u32; L.mo_adc.argtypes = [u64, C.c_uint8, C.POINTER(u64)]
        L.mo_adc_mul.restype = u32; L.mo_adc_mul.argtypes = [u64, u32, C.c_uint8, C.POINTER(u64)]
        L.mo_transform_size.restype = sz; L.mo_transform_size.argtypes = [u32]
        L.mo_widths.restype = C.c_int; L.mo_widths.argtypes = [sz, u32, vp]
        L.mo_weights.restype = C.c_int; L.mo_weights.argtypes = [sz, u32, vp, vp]
        L.mo_create.restype = vp; L.mo_create.argtypes = [u32, sz]
        L.mo_create_n.restype = vp; L.mo_create_n.argtypes = [u32, sz, sz]
        L.mo_destroy.restype = None; L.mo_destroy.argtypes = [vp]
        L.mo_size.restype = sz; L.mo_size.argtypes = [vp]
        for name, args in (("mo_set_u32", [vp, sz, u32]), ("mo_copy", [vp, sz, sz]),
                           ("mo_square_mul", [vp, sz, u32]), ("mo_set_multiplicand", [vp, sz, sz]),
                           ("mo_mul", [vp, sz, sz, u32]), ("mo_add", [vp, sz, sz]),
                           ("mo_sub_reg", [vp, sz, sz]), ("mo_sub_u32", [vp, sz, u32]),
                           ("mo_get_digits", [vp, sz, vp]), ("mo_set_digits", [vp, sz, vp]),
                           ("mo_get_raw", [vp, sz, vp]), ("mo_set_raw", [vp, sz, vp])):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = args
        _lib = L
    return _lib


_ref = None


def ref_lib():
    """The reference's own arith.h / ibdwt.h behind a C ABI (oracle/_ref); None if not built."""
    global _ref
    if _ref is None:
        build()
        if not os.path.exists(_REF_PATH):
            return None
        R = C.CDLL(_REF_PATH)
        u64, u32, sz, vp = C.c_uint64, C.c_uint32, C.c_size_t, C.c_void_p
        for name in ("ref_mod_add", "ref_mod_sub", "ref_mod_mul", "ref_mod_pow"):
            getattr(R, name).restype = u64
            getattr(R, name).argtypes = [u64, u64]
        for name in ("ref_mod_muli", "ref_mod_half", "ref_mod_invert", "ref_mod_root_nth"):
            getattr(R, name).restype = u64
            getattr(R, name).argtypes = [u64]
        R.ref_adc.restype = u32; R.ref_adc.argtypes = [u64, C.c_uint8, C.POINTER(u64)]
        R.ref_adc_mul.restype = u32; R.ref_adc_mul.argtypes = [u64, u32, C.c_uint8, C.POINTER(u64)]
        R.ref_transform_size.restype = sz; R.ref_transform_size.argtypes = [u32]
        R.ref_weights_widths.restype = C.c_int; R.ref_weights_widths.argtypes = [sz, u32, vp, vp]
        R.ref_roots.restype = None; R.ref_roots.argtypes = [sz, vp]
        R.ref_get_canonical.restype = None; R.ref_get_canonical.argtypes = [sz, vp, vp, vp, vp]
        _ref = R
    return _ref


def transform_size(p):
    return lib().mo_transform_size(p)


def widths(n, p):
    w = np.zeros(n, dtype=np.uint8)
    if lib().mo_widths(n, p, w.ctypes.data) != 0:
        raise ValueError("bad widths")
    return w


def weights(n, p):
    w = np.zeros(n, dtype=np.uint64)
    wi = np.zeros(n, dtype=np.uint64)
    if lib().mo_weights(n, p, w.ctypes.data, wi.ctypes.data) != 0:
        raise ValueError("n does not divide (P-1)/192")
    return w, wi


def digits_to_int(enc):
    """Canonical encoded digits (value | width << 32) -> Python int in [0, 2^p - 1]."""
    enc = np.asarray(enc, dtype=np.uint64)
    vals = (enc & np.uint64(0xFFFFFFFF)).astype(object)
    wd = (enc >> np.uint64(32)).astype(np.int64)
    pos = np.concatenate(([0], np.cumsum(wd)[:-1]))
    if len(enc) <= 1 << 15:
        return int(sum(int(v) << int(s) for v, s in zip(vals, pos)))
    # large: pack bits through numpy
    total = int(wd.sum())
    out = np.zeros((total + 63) // 64 + 1, dtype=np.uint64)
    v64 = enc & np.uint64(0xFFFFFFFF)
    word = (pos // 64).astype(np.int64)
    sh = (pos % 64).astype(np.uint64)
    lo = v64 << sh
    hi = np.where(sh > 0, v64 >> (np.uint64(64) - np.where(sh > 0, sh, np.uint64(1))), np.uint64(0))
    np.bitwise_or.at(out, word, lo)
    np.bitwise_or.at(out, word + 1, hi)
    return int.from_bytes(out.tobytes(), "little")


def int_to_digits(x, wd):
    """Python int -> encoded digits for the width table wd (value | width << 32)."""
    wd = np.asarray(wd, dtype=np.int64)
    n = len(wd)
    total = int(wd.sum())
    x %= (1 << total) - 1 if total > 0 else 1
    raw = np.frombuffer(x.to_bytes((total + 63) // 64 * 8 + 8, "little"), dtype=np.uint64)
    pos = np.concatenate(([0], np.cumsum(wd)[:-1]))
    word = (pos // 64).astype(np.int64)
    sh = (pos % 64).astype(np.uint64)
    lo = raw[word] >> sh
    hi = np.where(sh > 0, raw[word + 1] << (np.uint64(64) - np.where(sh > 0, sh, np.uint64(1))), np.uint64(0))
    mask = (np.uint64(1) << wd.astype(np.uint64)) - np.uint64(1)
    v = (lo | hi) & mask
    return v | (wd.astype(np.uint64) << np.uint64(32))


class OracleEngine:
    """CPU oracle with the reference's register API (engine.h:37-74)."""

    def __init__(self, p, regs, n=None):
        L = lib()
        self.p = p
        self._h = L.mo_create(p, regs) if n is None else L.mo_create_n(p, regs, n)
        if not self._h:
            raise ValueError(f"oracle: cannot create engine for p={p}")
        self.n = L.mo_size(self._h)
        self.regs = regs

    def close(self):
        if self._h:
            lib().mo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def get_size(self):
        return self.n

    def set(self, dst, a):
        lib().mo_set_u32(self._h, dst, a)

    def copy(self, dst, src):
        lib().mo_copy(self._h, dst, src)

    def square_mul(self, r, a=1):
        lib().mo_square_mul(self._h, r, a)

    def set_multiplicand(self, dst, src):
        lib().mo_set_multiplicand(self._h, dst, src)

    def mul(self, dst, src, a=1):
        lib().mo_mul(self._h, dst, src, a)

    def sub(self, r, a):
        lib().mo_sub_u32(self._h, r, a)

    def add(self, dst, src):
        lib().mo_add(self._h, dst, src)

    def sub_reg(self, dst, src):
        lib().mo_sub_reg(self._h, dst, src)

    def get_digits(self, src):
        out = np.zeros(self.n, dtype=np.uint64)
        lib().mo_get_digits(self._h, src, out.ctypes.data)
        return out

    def set_digits(self, dst, enc):
        enc = np.ascontiguousarray(enc, dtype=np.uint64)
        assert enc.shape == (self.n,)
        lib().mo_set_digits(self._h, dst, enc.ctypes.data)

    def widths(self):
        return widths(self.n, self.p)

    def get_int(self, src):
        """Value in [0, 2^p - 1) (2^p - 1 aliases to 0, engine.h:181-196)."""
        v = digits_to_int(self.get_digits(src))
        return 0 if v == (1 << self.p) - 1 else v

    def set_int(self, dst, x):
        self.set_digits(dst, int_to_digits(x % ((1 << self.p) - 1), self.widths()))
