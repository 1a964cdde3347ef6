"""Host-side mirror of the reference's register engine API over the C ABI of libmarin_b200.so.

`Engine` has the methods of `class engine` (include/marin/engine.h:16-303) with the same names,
argument meaning and error behaviour (errors raise, as the reference throws std::runtime_error;
the bool getters return False on a size mismatch).  `Digit` mirrors `engine::digit`
(engine.h:234-296).  There is no CPU fallback: if the CUDA library cannot be loaded or no B200 is
present, construction fails loudly.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

_lib = None

_SIGS = {
    "mb200_version": (C.c_char_p, []),
    "mb200_last_error": (C.c_char_p, []),
    "mb200_transform_size_for": (C.c_size_t, [C.c_uint32]),
    "mb200_create": (C.c_void_p, [C.c_uint32, C.c_size_t, C.c_uint32, C.c_int]),
    "mb200_create_split": (C.c_void_p, [C.c_uint32, C.c_size_t, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32]),
    "mb200_destroy": (None, [C.c_void_p]),
    "mb200_transform_size": (C.c_size_t, [C.c_void_p]),
    "mb200_register_count": (C.c_size_t, [C.c_void_p]),
    "mb200_max_multiplier": (C.c_uint32, [C.c_void_p]),
    "mb200_sync": (C.c_int, [C.c_void_p]),
    "mb200_set_u32": (C.c_int, [C.c_void_p, C.c_size_t, C.c_uint32]),
    "mb200_copy": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t]),
    "mb200_square_mul": (C.c_int, [C.c_void_p, C.c_size_t, C.c_uint32]),
    "mb200_set_multiplicand": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t]),
    "mb200_add_multiplicand": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t]),
    "mb200_mul": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint32]),
    "mb200_sub_u32": (C.c_int, [C.c_void_p, C.c_size_t, C.c_uint32]),
    "mb200_add": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t]),
    "mb200_sub_reg": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t]),
    "mb200_mul_add": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint32]),
    "mb200_addsub": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t]),
    "mb200_addsub_copy": (C.c_int, [C.c_void_p] + [C.c_size_t] * 6),
    "mb200_square_mul_copy": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint32]),
    "mb200_mul_copy": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint32]),
    "mb200_equal": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.POINTER(C.c_int)]),
    "mb200_get_digits": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "mb200_set_digits": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "mb200_register_data_size": (C.c_size_t, [C.c_void_p]),
    "mb200_get_raw": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "mb200_set_raw": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "mb200_checkpoint_size": (C.c_size_t, [C.c_void_p]),
    "mb200_get_checkpoint": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "mb200_set_checkpoint": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "mb200_bench_square_mul": (C.c_int, [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.POINTER(C.c_float)]),
    "mb200_square_mul_repeat": (C.c_int, [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32]),
    "mb200_bench_pass": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_uint32, C.POINTER(C.c_float)]),
    "mb200_is_single_cta": (C.c_int, [C.c_void_p]),
    "mb200_device_mem_info": (C.c_int, [C.c_uint32, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "mb200_create_batch": (C.c_void_p, [C.c_uint32, C.c_size_t, C.c_uint32, C.c_int, C.c_uint32]),
    "mb200_batch": (C.c_uint32, [C.c_void_p]),
    "mb200_get_digits_slab": (C.c_int, [C.c_void_p, C.c_uint32, C.c_size_t, C.c_void_p, C.c_size_t]),
    "mb200_set_digits_slab": (C.c_int, [C.c_void_p, C.c_uint32, C.c_size_t, C.c_void_p, C.c_size_t]),
    "mb200_graph_begin": (C.c_int, [C.c_void_p]),
    "mb200_graph_end": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "mb200_graph_launch": (C.c_int, [C.c_void_p, C.c_int]),
    "mb200_launch_count": (C.c_uint64, [C.c_void_p]),
    "mb200_host_alloc": (C.c_void_p, [C.c_size_t]),
    "mb200_host_free": (None, [C.c_void_p]),
}


def exported_symbols():
    """Every entry point include/marin_b200.h declares."""
    return sorted(_SIGS)


def load_library(build_if_needed=True):
    """dlopen libmarin_b200.so (building it first if stale); raises if it cannot be had."""
    global _lib
    if _lib is None:
        path = os.environ.get("MARIN_B200_LIB") or _build.LIB   # development aid: try another build of the library
        if build_if_needed and _build.nvcc_path() is not None and path == _build.LIB:
            path = _build.build()
        if not os.path.exists(path):
            raise RuntimeError("libmarin_b200.so is missing: run `python -m prmers_b200.build` (needs nvcc); "
                               "there is no CPU fallback")
        L = C.CDLL(path)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def transform_size(p):
    """ibdwt::transform_size(p) (ibdwt.h:17-43)."""
    return load_library().mb200_transform_size_for(p)


class MarinError(RuntimeError):
    pass


class Engine:
    """engine::create_gpu(p, reg_count, device, verbose) and the methods of `class engine`."""

    def __init__(self, p, reg_count, device=0, verbose=False, batch=1, force_S=0, force_R2=0):
        """batch > 1: `batch` independent register sets ("slabs") driven in lockstep by every call, one CTA per slab in
        each launch (single-CTA sizes only: many ECM curves at once); digits are read and written per slab.
        force_S / force_R2 (tests and tuning): row length of the contiguous pass / rows of a second strided pass."""
        L = load_library()
        self._L = L
        self.batch = batch
        if batch > 1:
            self._h = L.mb200_create_batch(p, reg_count, device, 1 if verbose else 0, batch)
        elif force_S or force_R2:
            self._h = L.mb200_create_split(p, reg_count, device, 1 if verbose else 0, force_S, force_R2)
        else:
            self._h = L.mb200_create(p, reg_count, device, 1 if verbose else 0)
        if not self._h:
            raise MarinError(L.mb200_last_error().decode())
        self.p = p
        self.n = L.mb200_transform_size(self._h)
        self.reg_count = reg_count
        self.device = device

    # -- lifetime -------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._L.mb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, ok):
        if not ok:
            raise MarinError(self._L.mb200_last_error().decode())

    # -- engine.h:40-74 -------------------------------------------------------------------------
    def get_size(self):
        return self.n

    def sync(self):
        self._ck(self._L.mb200_sync(self._h))

    def set(self, dst, a):
        self._ck(self._L.mb200_set_u32(self._h, dst, a))

    def copy(self, dst, src):
        self._ck(self._L.mb200_copy(self._h, dst, src))

    def square_mul(self, r, a=1):
        self._ck(self._L.mb200_square_mul(self._h, r, a))

    def set_multiplicand(self, dst, src):
        self._ck(self._L.mb200_set_multiplicand(self._h, dst, src))

    def add_multiplicand(self, dst, a, b):
        """dst <- multiplicand of (x + y) from the multiplicands a of x and b of y (needs max_multiplier() >= 2)."""
        self._ck(self._L.mb200_add_multiplicand(self._h, dst, a, b))

    def set_multiplicand2(self, dst, src):
        self.set_multiplicand(dst, src)

    def mul(self, dst, src, a=1):
        self._ck(self._L.mb200_mul(self._h, dst, src, a))

    def mul_new(self, dst, src, a=1):
        self.mul(dst, src, a)

    def sub(self, r, a):
        self._ck(self._L.mb200_sub_u32(self._h, r, a))

    def add(self, dst, src):
        self._ck(self._L.mb200_add(self._h, dst, src))

    def sub_reg(self, dst, src):
        self._ck(self._L.mb200_sub_reg(self._h, dst, src))

    # -- engine.h:65-131 fused helpers ----------------------------------------------------------
    def mul_add(self, dst, mul_src, add_src, a=1):
        self._ck(self._L.mb200_mul_add(self._h, dst, mul_src, add_src, a))

    def addsub(self, sum_out, diff_out, a, b):
        self._ck(self._L.mb200_addsub(self._h, sum_out, diff_out, a, b))

    def addsub_copy(self, s, d, s_copy, d_copy, a, b):
        self._ck(self._L.mb200_addsub_copy(self._h, s, d, s_copy, d_copy, a, b))

    def square_mul_copy(self, r, dst_copy, a=1):
        self._ck(self._L.mb200_square_mul_copy(self._h, r, dst_copy, a))

    def mul_copy(self, dst, src, dst_copy, a=1):
        self._ck(self._L.mb200_mul_copy(self._h, dst, src, dst_copy, a))

    def mul_pair_unit(self, dst0, src0, dst1, src1):
        self.set_multiplicand(src0, src0); self.mul(dst0, src0)
        self.set_multiplicand(src1, src1); self.mul(dst1, src1)

    def mul_pair_prepared(self, dst0, m0, dst1, m1, a0=1, a1=1):
        self.mul(dst0, m0, a0); self.mul(dst1, m1, a1)

    def xdbl_tail_uv(self, x_out, z_out, u_work, v_reg, a24_mul, tmp_e_mul, tmp_v_mul):
        self.sub_reg(u_work, v_reg)
        self.set_multiplicand(tmp_v_mul, v_reg)
        self.mul(x_out, tmp_v_mul)
        self.set_multiplicand(tmp_e_mul, u_work)
        self.mul_add(u_work, a24_mul, v_reg)
        self.mul_copy(u_work, tmp_e_mul, z_out)

    def pow(self, dst, src, e):
        """dst = src^e, src is erased (engine.h:160-170)."""
        self.set_multiplicand(src, src)
        self.set(dst, 1)
        if e == 0:
            return
        for i in range(int(e).bit_length() - 1, -1, -1):
            self.square_mul(dst)
            if (e >> i) & 1:
                self.mul(dst, src)

    def is_equal(self, lhs, rhs):
        out = C.c_int(0)
        self._ck(self._L.mb200_equal(self._h, lhs, rhs, C.byref(out)))
        return bool(out.value)

    # -- digits / big integers (engine.h:24-25,173-232) -----------------------------------------
    def get(self, src, out=None, slab=0):
        """Canonical encoded digits (value | width << 32) as a uint64 array."""
        if out is None:
            out = np.empty(self.n, dtype=np.uint64)
        if slab:
            self._ck(self._L.mb200_get_digits_slab(self._h, slab, src, out.ctypes.data, self.n))
        else:
            self._ck(self._L.mb200_get_digits(self._h, src, out.ctypes.data, self.n))
        return out

    def set_digits(self, dst, enc, slab=0):
        enc = np.ascontiguousarray(enc, dtype=np.uint64)
        if enc.shape != (self.n,):
            raise MarinError("digit count must equal the transform size")
        if slab:
            self._ck(self._L.mb200_set_digits_slab(self._h, slab, dst, enc.ctypes.data, self.n))
        else:
            self._ck(self._L.mb200_set_digits(self._h, dst, enc.ctypes.data, self.n))

    def widths(self):
        return _widths(self.n, self.p)

    def get_mpz(self, src, slab=0):
        """Value in [0, 2^p - 1): all-ones (= 2^p - 1) maps to 0 (engine.h:173-203)."""
        v = digits_to_int(self.get(src, slab=slab))
        return 0 if v == (1 << self.p) - 1 else v

    def set_mpz(self, dst, z, slab=0):
        """Reduces z modulo 2^p - 1 first (the reference's default does not, engine.h:205-232)."""
        self.set_digits(dst, int_to_digits(int(z) % ((1 << self.p) - 1), _widths(self.n, self.p)), slab=slab)

    def slab(self, index):
        """A per-slab view for the host side of a batched run: digits in and out of ONE slab."""
        return SlabView(self, index)

    def digit(self, src):
        return Digit(self, src)

    # -- raw images (engine.h:133-146) ----------------------------------------------------------
    def get_register_data_size(self):
        return self._L.mb200_register_data_size(self._h)

    def get_data(self, data, src):
        if len(data) != self.get_register_data_size():
            return False
        buf = (C.c_char * len(data)).from_buffer(data)
        self._ck(self._L.mb200_get_raw(self._h, src, C.addressof(buf), len(data)))
        return True

    def set_data(self, dst, data):
        if len(data) != self.get_register_data_size():
            return False
        buf = (C.c_char * len(data)).from_buffer_copy(bytes(data))
        self._ck(self._L.mb200_set_raw(self._h, dst, C.addressof(buf), len(data)))
        return True

    def get_checkpoint_size(self):
        return self._L.mb200_checkpoint_size(self._h)

    def get_checkpoint(self, data):
        if len(data) != self.get_checkpoint_size():
            return False
        buf = (C.c_char * len(data)).from_buffer(data)
        self._ck(self._L.mb200_get_checkpoint(self._h, C.addressof(buf), len(data)))
        return True

    def set_checkpoint(self, data):
        if len(data) != self.get_checkpoint_size():
            return False
        buf = (C.c_char * len(data)).from_buffer_copy(bytes(data))
        self._ck(self._L.mb200_set_checkpoint(self._h, C.addressof(buf), len(data)))
        return True

    def is_aevum_backend(self):
        return False

    def release_gpu_resources_for_lowmem_handoff(self):
        pass

    # -- measurement helpers --------------------------------------------------------------------
    def max_multiplier(self):
        return self._L.mb200_max_multiplier(self._h)

    def bench_square_mul(self, r, count, a=1):
        """`count` square_mul(r, a) timed with CUDA events on the engine's stream; returns ms."""
        ms = C.c_float(0)
        self._ck(self._L.mb200_bench_square_mul(self._h, r, a, count, C.byref(ms)))
        return ms.value

    def bench_pass(self, r, which, count):
        """One kernel of square_mul alone (0 strided fwd, 1 middle, 2 strided inv + carry, 3 fix-up); ms."""
        ms = C.c_float(0)
        self._ck(self._L.mb200_bench_pass(self._h, r, which, count, C.byref(ms)))
        return ms.value

    def is_single_cta(self):
        return bool(self._L.mb200_is_single_cta(self._h))

    # -- record a fixed op sequence, replay it as one kernel launch (single-CTA engines; include/marin_b200.h) ---
    def graph_begin(self):
        self._ck(self._L.mb200_graph_begin(self._h))

    def graph_end(self):
        gid = C.c_int(-1)
        self._ck(self._L.mb200_graph_end(self._h, C.byref(gid)))
        return gid.value

    def graph_launch(self, gid):
        self._ck(self._L.mb200_graph_launch(self._h, gid))

    def square_mul_repeat(self, r, count, a=1):
        self._ck(self._L.mb200_square_mul_repeat(self._h, r, a, count))

    def launch_count(self):
        return self._L.mb200_launch_count(self._h)


class SlabView:
    """One slab of a batched engine as seen by host code that reads / writes residues (get, get_mpz, set_mpz, set_digits);
    the arithmetic itself is issued on the engine, for all slabs at once."""

    def __init__(self, eng, index):
        self.eng, self.index, self.p, self.n = eng, index, eng.p, eng.n

    def get_size(self): return self.n
    def get(self, src, out=None): return self.eng.get(src, out=out, slab=self.index)
    def get_mpz(self, src): return self.eng.get_mpz(src, slab=self.index)
    def set_mpz(self, dst, z): self.eng.set_mpz(dst, z, slab=self.index)
    def set_digits(self, dst, enc): self.eng.set_digits(dst, enc, slab=self.index)


def device_mem_info(device=0):
    """(free, total) bytes of device memory (cudaMemGetInfo): the P-1 stage-2 planner's budget."""
    L = load_library()
    free, total = C.c_size_t(0), C.c_size_t(0)
    if not L.mb200_device_mem_info(device, C.byref(free), C.byref(total)):
        raise MarinError(L.mb200_last_error().decode())
    return free.value, total.value


def create_gpu(p, reg_count, device=0, verbose=False):
    """engine::create_gpu (engine.h:301, src/marin/gpu.cpp:52-150): single backend, no dispatch."""
    return Engine(p, reg_count, device, verbose)


class Digit:
    """engine::digit (engine.h:234-296)."""

    def __init__(self, eng, src):
        self._d = eng.get(src)

    def get_size(self):
        return len(self._d)

    def val(self, i):
        return int(self._d[i]) & 0xFFFFFFFF

    def width(self, i):
        return int(self._d[i]) >> 32

    def res64(self):
        r, s = 0, 0
        for d in self._d[:80]:
            d = int(d)
            r += (d & 0xFFFFFFFF) << s
            s += d >> 32
            if s >= 64:
                break
        return r & 0xFFFFFFFFFFFFFFFF

    def equal_to(self, a):
        v = self._d & np.uint64(0xFFFFFFFF)
        r = int(a)
        for k in range(min(len(self._d), 80)):
            w = int(self._d[k]) >> 32
            if (r & ((1 << w) - 1)) != int(v[k]):
                return False
            r >>= w
        return r == 0 and not np.any(v[80:]) if len(v) > 80 else r == 0

    def equal_to_Mp(self):
        v = self._d & np.uint64(0xFFFFFFFF)
        w = self._d >> np.uint64(32)
        return bool(np.all(v == (np.uint64(1) << w) - np.uint64(1)))

    def to_int(self):
        return digits_to_int(self._d)


# ---- digit packing (engine.h:173-232, AlgoUtils.hpp:165-187) ------------------------------------

_width_cache = {}


def _widths(n, p):
    """width[j] = ceil(p (j+1) / n) - ceil(p j / n) (ibdwt.h:122-133)."""
    key = (n, p)
    if key not in _width_cache:
        j = np.arange(n + 1, dtype=np.uint64)
        c = (np.uint64(p) * j + np.uint64(n - 1)) // np.uint64(n)
        _width_cache[key] = (c[1:] - c[:-1]).astype(np.uint8)
    return _width_cache[key]


def digits_to_int(enc):
    enc = np.asarray(enc, dtype=np.uint64)
    wd = (enc >> np.uint64(32)).astype(np.int64)
    pos = np.concatenate(([0], np.cumsum(wd)[:-1]))
    total = int(wd.sum())
    out = np.zeros((total + 63) // 64 + 2, dtype=np.uint64)
    v64 = enc & np.uint64(0xFFFFFFFF)
    word = (pos // 64).astype(np.int64)
    sh = (pos % 64).astype(np.uint64)
    lo = v64 << sh
    hi = np.where(sh > 0, v64 >> (np.uint64(64) - np.where(sh > 0, sh, np.uint64(1))), np.uint64(0))
    # digits never overlap bit-wise when canonical, but a digit may straddle two words
    np.add.at(out, word, lo)
    np.add.at(out, word + 1, hi)
    return int.from_bytes(out.tobytes(), "little")


def int_to_digits(x, wd):
    wd = np.asarray(wd, dtype=np.int64)
    total = int(wd.sum())
    raw = np.frombuffer(int(x).to_bytes((total + 63) // 64 * 8 + 8, "little"), dtype=np.uint64)
    pos = np.concatenate(([0], np.cumsum(wd)[:-1]))
    word = (pos // 64).astype(np.int64)
    sh = (pos % 64).astype(np.uint64)
    lo = raw[word] >> sh
    hi = np.where(sh > 0, raw[word + 1] << (np.uint64(64) - np.where(sh > 0, sh, np.uint64(1))), np.uint64(0))
    mask = (np.uint64(1) << wd.astype(np.uint64)) - np.uint64(1)
    return ((lo | hi) & mask) | (wd.astype(np.uint64) << np.uint64(32))
