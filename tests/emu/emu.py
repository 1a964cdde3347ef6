"""ctypes wrapper of tests/emu/libmarin_emu.so (TEST INFRASTRUCTURE: CPU emulation of the CUDA kernels)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_EXTRA = os.environ.get("MARIN_EMU_CFLAGS", "").split()          # e.g. -DMARIN_TILE_C=4
_SO = os.path.join(_HERE, "libmarin_emu" + ("_" + "_".join(f.strip("-").replace("=", "") for f in _EXTRA) if _EXTRA else "") + ".so")
_SRC = os.path.join(_HERE, "emu_engine.cpp")
_CSRC = os.path.join(_HERE, "..", "..", "prmers_b200", "csrc")


def build():
    deps = [_SRC] + [os.path.join(_CSRC, f) for f in ("gl64.cuh", "marin_kernels.cuh", "marin_plan.hpp", "marin_engine_core.hpp")]
    if os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(d) for d in deps):
        return
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-std=c++17", "-O2", "-fopenmp", "-fPIC", "-shared"] + _EXTRA + ["-o", _SO, _SRC]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        r = subprocess.run([c for c in cmd if c != "-fopenmp"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(r.stderr)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, sz, u32 = C.c_void_p, C.c_size_t, C.c_uint32
        L.emu_last_error.restype = C.c_char_p
        L.emu_create.restype = vp; L.emu_create.argtypes = [u32, sz, u32, u32]
        L.emu_create_batch.restype = vp; L.emu_create_batch.argtypes = [u32, sz, u32]
        L.emu_get_digits_slab.restype = C.c_int; L.emu_get_digits_slab.argtypes = [vp, u32, sz, vp]
        L.emu_set_digits_slab.restype = C.c_int; L.emu_set_digits_slab.argtypes = [vp, u32, sz, vp]
        L.emu_destroy.argtypes = [vp]
        L.emu_size.restype = sz; L.emu_size.argtypes = [vp]
        for f in ("emu_is_small", "emu_plan_R", "emu_plan_S", "emu_plan_R2", "emu_a_max"):
            getattr(L, f).argtypes = [vp]; getattr(L, f).restype = u32
        for f, a in (("emu_set_u32", [vp, sz, u32]), ("emu_copy", [vp, sz, sz]), ("emu_square_mul", [vp, sz, u32]),
                     ("emu_set_multiplicand", [vp, sz, sz]), ("emu_mul", [vp, sz, sz, u32]), ("emu_add_multiplicand", [vp, sz, sz, sz]), ("emu_sub_u32", [vp, sz, u32]),
                     ("emu_add", [vp, sz, sz]), ("emu_sub_reg", [vp, sz, sz]), ("emu_get_digits", [vp, sz, vp]),
                     ("emu_set_digits", [vp, sz, vp]), ("emu_equal", [vp, sz, sz, vp]), ("emu_graph_begin", [vp]),
                     ("emu_graph_end", [vp, vp]), ("emu_graph_launch", [vp, C.c_int]), ("emu_graph_shape", [vp, C.c_int, vp, vp]),
                     ("emu_addsub", [vp, sz, sz, sz, sz]), ("emu_addsub_copy", [vp, sz, sz, sz, sz, sz, sz]),
                     ("emu_square_mul_copy", [vp, sz, sz, u32]), ("emu_mul_copy", [vp, sz, sz, sz, u32]),
                     ("emu_mul_add", [vp, sz, sz, sz, u32])):
            getattr(L, f).argtypes = a; getattr(L, f).restype = C.c_int
        L.emu_transform_size.restype = sz; L.emu_transform_size.argtypes = [u32]
        L.emu_checkpoint_size.restype = sz; L.emu_checkpoint_size.argtypes = [vp]
        L.emu_get_checkpoint.restype = C.c_int; L.emu_get_checkpoint.argtypes = [vp, vp]
        L.emu_set_checkpoint.restype = C.c_int; L.emu_set_checkpoint.argtypes = [vp, vp]
        _lib = L
    return _lib


class EmuEngine:
    def __init__(self, p, regs, force_S=0, force_R2=0, batch=1):
        L = lib()
        self.batch = batch
        self._h = L.emu_create_batch(p, regs, batch) if batch > 1 else L.emu_create(p, regs, force_S, force_R2)
        if not self._h:
            raise ValueError(L.emu_last_error().decode())
        self.p, self.n, self.reg_count = p, L.emu_size(self._h), regs
        self.small = bool(L.emu_is_small(self._h))
        self.R, self.S, self.R2 = L.emu_plan_R(self._h), L.emu_plan_S(self._h), L.emu_plan_R2(self._h)
        self.a_max = L.emu_a_max(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().emu_destroy(self._h); self._h = None

    def _ck(self, ok):
        if not ok:
            raise RuntimeError(lib().emu_last_error().decode())

    def set(self, d, a): self._ck(lib().emu_set_u32(self._h, d, a))
    def copy(self, d, s): self._ck(lib().emu_copy(self._h, d, s))
    def square_mul(self, r, a=1): self._ck(lib().emu_square_mul(self._h, r, a))
    def set_multiplicand(self, d, s): self._ck(lib().emu_set_multiplicand(self._h, d, s))
    def mul(self, d, s, a=1): self._ck(lib().emu_mul(self._h, d, s, a))
    def add_multiplicand(self, d, a, b): self._ck(lib().emu_add_multiplicand(self._h, d, a, b))
    def max_multiplier(self): return self.a_max
    def sub(self, r, a): self._ck(lib().emu_sub_u32(self._h, r, a))
    def add(self, d, s): self._ck(lib().emu_add(self._h, d, s))
    def sub_reg(self, d, s): self._ck(lib().emu_sub_reg(self._h, d, s))
    def addsub(self, s, d, a, b): self._ck(lib().emu_addsub(self._h, s, d, a, b))
    def addsub_copy(self, s, d, sc, dc, a, b): self._ck(lib().emu_addsub_copy(self._h, s, d, sc, dc, a, b))
    def square_mul_copy(self, r, c, a=1): self._ck(lib().emu_square_mul_copy(self._h, r, c, a))
    def mul_add(self, d, ms, as_, a=1): self._ck(lib().emu_mul_add(self._h, d, ms, as_, a))
    def mul_copy(self, d, s, c, a=1): self._ck(lib().emu_mul_copy(self._h, d, s, c, a))

    # recorded op sequences (the emulated k_program)
    def is_single_cta(self): return self.small
    def graph_begin(self): self._ck(lib().emu_graph_begin(self._h))

    def graph_end(self):
        gid = C.c_int(-1)
        self._ck(lib().emu_graph_end(self._h, C.byref(gid)))
        return gid.value

    def graph_launch(self, gid): self._ck(lib().emu_graph_launch(self._h, gid))

    def graph_shape(self, gid):
        """(length of the stored schedule in the scheduler's cost units of 0.1 us, number of ops) of a recorded program."""
        lv, ops = C.c_int(0), C.c_int(0)
        self._ck(lib().emu_graph_shape(self._h, gid, C.byref(lv), C.byref(ops)))
        return lv.value, ops.value

    def get_checkpoint_size(self): return lib().emu_checkpoint_size(self._h)

    def get_checkpoint(self, data):
        if len(data) != self.get_checkpoint_size():
            return False
        buf = (C.c_char * len(data)).from_buffer(data)
        self._ck(lib().emu_get_checkpoint(self._h, C.addressof(buf)))
        return True

    def set_checkpoint(self, data):
        if len(data) != self.get_checkpoint_size():
            return False
        buf = (C.c_char * len(data)).from_buffer_copy(bytes(data))
        self._ck(lib().emu_set_checkpoint(self._h, C.addressof(buf)))
        return True

    def is_equal(self, l, r):
        out = C.c_int(0)
        self._ck(lib().emu_equal(self._h, l, r, C.byref(out)))
        return bool(out.value)

    def get_digits(self, s, slab=0):
        out = np.zeros(self.n, dtype=np.uint64)
        self._ck(lib().emu_get_digits_slab(self._h, slab, s, out.ctypes.data))
        return out

    def get(self, s, out=None, slab=0):
        return self.get_digits(s, slab)

    def get_mpz(self, s, slab=0):
        from oracle import marin_oracle as mo
        v = mo.digits_to_int(self.get_digits(s, slab))
        return 0 if v == (1 << self.p) - 1 else v

    def set_mpz(self, d, z, slab=0):
        from oracle import marin_oracle as mo
        self.set_digits(d, mo.int_to_digits(int(z) % ((1 << self.p) - 1), mo.widths(self.n, self.p)), slab)

    def slab(self, index):
        from prmers_b200.engine import SlabView
        return SlabView(self, index)

    def get_size(self):
        return self.n

    def set_digits(self, d, enc, slab=0):
        enc = np.ascontiguousarray(enc, dtype=np.uint64)
        self._ck(lib().emu_set_digits_slab(self._h, slab, d, enc.ctypes.data))
