"""Checkpoint files for the mode loops, in the reference's container format with a NEW backend tag.

Layout (RunPrpOrLlMarin.cpp:153-211): int32 version = 2, uint32 p, uint32 mode (1 = PRP, 2 = LL),
uint32 backend, uint32 iteration, float64 elapsed seconds, the raw register image
(engine::get_checkpoint), uint32 CRC.  The CRC is the running CRC-32 of everything before it,
stored as ~crc ^ 0xa23777ac (include/marin/file.h:48-110).  Backend tags: 1 = Marin (OpenCL),
2 = Aevum, **3 = this engine**: raw images are layout specific, so a file written by another backend
is ignored rather than loaded (the reference does the same for its two backends).
Rotation: write `<file>.new`, move the current file to `<file>.old`, move `.new` into place.

The other mode loops keep the reference's own field orders (all little endian, CRC word last):
  LL-safe  (RunLlSafeMarin.cpp:418-477): int32 version = 3, uint32 p, uint32 backend, uint64 iterations done, float64
           elapsed, uint64 itersave, uint64 jsave, raw image.
  P-1 stage 1 (RunPM1.cpp:6219-6267): int32 version = 3, uint32 p, uint32 bit index, float64 elapsed, raw image, uint64
           checkpass, blocks since check, bits in block, current block length, uint8 in_lot, eacc and wbits as
           length-prefixed hex strings, uint64 chunk index, start prime, uint8 first, uint64 processed bits, bits in
           chunk; the backend lives in a side file `<file>.backend` (RunPM1.cpp:154-181), here "marin_b200".
  ECM stage 1 (RunEcmTwistedEdwards.cpp:1866-1930): int32 version, uint32 p, uint32 step, uint32 bits of K, uint64 B1,
           float64 elapsed, uint64 curve seed, uint8 curve family, raw image.  The reference's file has no backend
           field, so the version word carries it here: 0x0301 (the reference accepts 1 only and skips the file).
  ECM stage 2 (RunEcmTwistedEdwards.cpp:1940-2068, version 6): int32 version, uint32 p, uint32 index of the next stage-2
           prime, uint32 number of stage-2 primes, uint64 B1, uint64 B2, uint64 curve seed, uint8 curve family, float64
           elapsed, uint8 in_chunk, uint32 chunk start index, chunk end index, chunk bits, steps done inside the chunk,
           uint8 has_base (+ six length-prefixed hex strings: X, Y, T of the affine base and of its negative), raw image.
           Version word 0x0306 for the same reason (the reference accepts 2 ... 6).
"""
import os
import struct
import zlib

BACKEND_TAG = 3
MODE_PRP, MODE_LL = 1, 2
_MASK = 0xA23777AC
_HDR = struct.Struct("<iIIIId")


def checkpoint_name(p, mode=MODE_PRP):
    return ("llunsafe_" if mode == MODE_LL else "") + f"m_{p}.ckpt"


def _crc_word(crc):
    return (~crc & 0xFFFFFFFF) ^ _MASK


def save(path, eng, p, iteration, elapsed=0.0, mode=MODE_PRP):
    data = bytearray(eng.get_checkpoint_size())
    if not eng.get_checkpoint(data):
        raise RuntimeError("engine refused the checkpoint buffer")
    head = _HDR.pack(2, p, mode, BACKEND_TAG, iteration, float(elapsed))
    crc = zlib.crc32(bytes(data), zlib.crc32(head))
    new, old = path + ".new", path + ".old"
    with open(new, "wb") as f:
        f.write(head); f.write(data); f.write(struct.pack("<I", _crc_word(crc)))
    if os.path.exists(old):
        os.remove(old)
    if os.path.exists(path):
        os.rename(path, old)
    os.rename(new, path)


def load(path, eng, p, mode=MODE_PRP):
    """Returns (iteration, elapsed) and restores the registers, or None if the file is missing, foreign
    (other exponent / mode / backend) or damaged; nothing is loaded in that case."""
    try:
        with open(path, "rb") as f:
            blob = f.read()
    except OSError:
        return None
    size = eng.get_checkpoint_size()
    if len(blob) != _HDR.size + size + 4:
        return None
    version, rp, rmode, rbackend, it, et = _HDR.unpack_from(blob)
    if version != 2 or rp != p or rmode != mode or rbackend != BACKEND_TAG:
        return None
    (stored,) = struct.unpack_from("<I", blob, _HDR.size + size)
    if stored != _crc_word(zlib.crc32(blob[:_HDR.size + size])):
        return None
    if not eng.set_checkpoint(bytearray(blob[_HDR.size:_HDR.size + size])):
        return None
    return it, et


def load_latest(path, eng, p, mode=MODE_PRP):
    """Current file first, then `.old` (RunPrpOrLlMarin.cpp:215-216)."""
    return load(path, eng, p, mode) or load(path + ".old", eng, p, mode)


# ---- the other mode loops --------------------------------------------------------------------------------------------

PM1_BACKEND_NAME = "marin_b200"
ECM_VERSION = 0x0301
ECM_FAMILY = {None: 0, "none": 0, "16": 1, "iv163": 2}


def _raw(eng):
    data = bytearray(eng.get_checkpoint_size())
    if not eng.get_checkpoint(data):
        raise RuntimeError("engine refused the checkpoint buffer")
    return bytes(data)


def _write_rotating(path, body, keep_old=True):
    """body + CRC word into `<path>.new`, current file to `.old`, `.new` into place."""
    new, old = path + ".new", path + ".old"
    with open(new, "wb") as f:
        f.write(body); f.write(struct.pack("<I", _crc_word(zlib.crc32(body))))
    if os.path.exists(old):
        os.remove(old)
    if os.path.exists(path):
        os.rename(path, old)
    os.rename(new, path)
    if not keep_old and os.path.exists(old):
        os.remove(old)


def _read_verified(path):
    """File content without its CRC word, or None when missing or damaged."""
    try:
        with open(path, "rb") as f:
            blob = f.read()
    except OSError:
        return None
    if len(blob) < 8:
        return None
    (stored,) = struct.unpack_from("<I", blob, len(blob) - 4)
    return blob[:-4] if stored == _crc_word(zlib.crc32(blob[:-4])) else None


_LLS = struct.Struct("<iIIQdQQ")


def llsafe_checkpoint_name(p):
    return f"llsafe_m_{p}.ckpt"


def save_llsafe(path, eng, p, iter_done, itersave, jsave, elapsed=0.0):
    _write_rotating(path, _LLS.pack(3, p, BACKEND_TAG, iter_done, float(elapsed), itersave, jsave) + _raw(eng))


def load_llsafe(path, eng, p):
    """(iter_done, itersave, jsave, elapsed) with the registers restored, or None (missing, damaged, foreign)."""
    for f in (path, path + ".old"):
        body = _read_verified(f)
        if body is None or len(body) != _LLS.size + eng.get_checkpoint_size():
            continue
        version, rp, backend, it, et, isave, jsave = _LLS.unpack_from(body)
        if version != 3 or rp != p or backend != BACKEND_TAG:
            continue
        if eng.set_checkpoint(bytearray(body[_LLS.size:])):
            return it, isave, jsave, et
    return None


def pm1_checkpoint_name(p):
    return f"pm1_m_{p}.ckpt"


def _hexfield(v):
    h = format(int(v), "x").encode()
    return struct.pack("<I", len(h)) + h


def save_pm1(path, eng, p, bit_index, state, elapsed=0.0):
    """`state`: checkpass, blocks_since_check, bits_in_block, current_block_len, in_lot, eacc, wbits, processed_bits, bits."""
    body = struct.pack("<iIId", 3, p, bit_index, float(elapsed)) + _raw(eng)
    body += struct.pack("<QQQQB", state["checkpass"], state["blocks_since_check"], state["bits_in_block"], state["current_block_len"], 1 if state["in_lot"] else 0)
    body += _hexfield(state["eacc"]) + _hexfield(state["wbits"])
    body += struct.pack("<QQBQQ", 0, 0, 1, state["processed_bits"], state["bits"])
    _write_rotating(path, body, keep_old=False)          # the reference removes `.old` after the rotation (RunPM1.cpp:6222)
    with open(path + ".backend.new", "w") as f:
        f.write(PM1_BACKEND_NAME + "\n")
    os.replace(path + ".backend.new", path + ".backend")


def load_pm1(path, eng, p):
    """(bit_index, state, elapsed) with the registers restored, or None.  A file without the side file that names this
    backend is not loaded (the reference reads a missing side file as "Marin OpenCL registers")."""
    try:
        with open(path + ".backend") as f:
            if f.read().split()[:1] != [PM1_BACKEND_NAME]:
                return None
    except OSError:
        return None
    body = _read_verified(path)
    size = eng.get_checkpoint_size()
    if body is None or len(body) < 20 + size + 33:
        return None
    version, rp, i, et = struct.unpack_from("<iIId", body)
    if version != 3 or rp != p:
        return None
    o = 20 + size
    chk, blks, bib, cbl, inlot = struct.unpack_from("<QQQQB", body, o); o += 33
    vals = []
    for _ in range(2):
        (ln,) = struct.unpack_from("<I", body, o); o += 4
        vals.append(int(body[o:o + ln] or b"0", 16)); o += ln
    _, _, _, processed, bits = struct.unpack_from("<QQBQQ", body, o)
    if not eng.set_checkpoint(bytearray(body[20:20 + size])):
        return None
    return i, {"checkpass": chk, "blocks_since_check": blks, "bits_in_block": bib, "current_block_len": cbl, "in_lot": bool(inlot),
               "eacc": vals[0], "wbits": vals[1], "processed_bits": processed, "bits": bits}, et


_ECM = struct.Struct("<iIIIQdQB")


def ecm_checkpoint_name(p, c):
    return f"ecm_te_m_{p}_c{c}.ckpt"


def save_ecm(path, eng, p, step, k_bits, b1, curve_seed, family=None, elapsed=0.0):
    _write_rotating(path, _ECM.pack(ECM_VERSION, p, step, k_bits, b1, float(elapsed), curve_seed, ECM_FAMILY[family]) + _raw(eng), keep_old=False)


def load_ecm(path, eng, p, k_bits, b1, curve_seed, family=None):
    """Step count of the saved stage-1 state (registers restored), or None: other exponent, bound, curve or backend."""
    for f in (path, path + ".old"):
        body = _read_verified(f)
        if body is None or len(body) != _ECM.size + eng.get_checkpoint_size():
            continue
        version, rp, step, nbb, rb1, et, seed, fam = _ECM.unpack_from(body)
        if version != ECM_VERSION or rp != p or nbb != k_bits or rb1 != b1 or seed != curve_seed or fam != ECM_FAMILY[family]:
            continue
        if eng.set_checkpoint(bytearray(body[_ECM.size:])):
            return step
    return None


ECM2_VERSION = 0x0306
_ECM2 = struct.Struct("<iIIIQQQBdBIIIIB")


def ecm2_checkpoint_name(p, c):
    return f"ecm2_te_m_{p}_c{c}.ckpt"


def save_ecm2(path, eng, p, idx, count, b1, b2, curve_seed, family=None, elapsed=0.0, chunk=None, base=None):
    """Stage-2 state: `idx` = index of the next stage-2 prime (the start of the chunk at work when `chunk` is given),
    `chunk` = (start index, end index, bits of the chunk scalar, steps done) when the scalar multiplication of a chunk was
    interrupted, `base` = the six residues X, Y, T, -X, Y, -T of the affine base of that chunk (save_ckpt2_ex)."""
    c = chunk or (0, 0, 0, 0)
    head = _ECM2.pack(ECM2_VERSION, p, idx, count, b1, b2, curve_seed, ECM_FAMILY[family], float(elapsed),
                      1 if chunk else 0, c[0], c[1], c[2], c[3], 1 if base else 0)
    blobs = b"".join(_hexfield(v) for v in (base or ()))
    _write_rotating(path, head + blobs + _raw(eng), keep_old=True)


def load_ecm2(path, eng, p, count, b1, b2, curve_seed, family=None):
    """{"idx", "elapsed", "chunk" (or None), "base" (or None)} with the registers restored, or None: other exponent,
    bounds, prime count, curve, family or backend (read_ckpt2_one; the file, then `.old`)."""
    size = eng.get_checkpoint_size()
    for f in (path, path + ".old"):
        body = _read_verified(f)
        if body is None or len(body) < _ECM2.size + size:
            continue
        version, rp, idx, cnt, rb1, rb2, seed, fam, et, in_chunk, cs, ce, cb, cd, has_base = _ECM2.unpack_from(body)
        if version != ECM2_VERSION or rp != p or cnt != count or rb1 != b1 or rb2 != b2 or seed != curve_seed or fam != ECM_FAMILY[family]:
            continue
        o, base = _ECM2.size, None
        if has_base:
            base = []
            for _ in range(6):
                (ln,) = struct.unpack_from("<I", body, o); o += 4
                base.append(int(body[o:o + ln] or b"0", 16)); o += ln
        if len(body) != o + size or not eng.set_checkpoint(bytearray(body[o:o + size])):
            continue
        return {"idx": idx, "elapsed": et, "chunk": (cs, ce, cb, cd) if in_chunk else None, "base": base}
    return None


def remove_ecm2(path):
    """A finished stage 2 leaves no file behind (:3838-3846)."""
    for f in (path, path + ".old", path + ".new"):
        if os.path.exists(f):
            os.remove(f)
