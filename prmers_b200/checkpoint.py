"""Checkpoint files for the mode loops, in the reference's container format with a NEW backend tag.

Layout (RunPrpOrLlMarin.cpp:153-211): int32 version = 2, uint32 p, uint32 mode (1 = PRP, 2 = LL),
uint32 backend, uint32 iteration, float64 elapsed seconds, the raw register image
(engine::get_checkpoint), uint32 CRC.  The CRC is the running CRC-32 of everything before it,
stored as ~crc ^ 0xa23777ac (include/marin/file.h:48-110).  Backend tags: 1 = Marin (OpenCL),
2 = Aevum, **3 = this engine**: raw images are layout specific, so a file written by another backend
is ignored rather than loaded (the reference does the same for its two backends).
Rotation: write `<file>.new`, move the current file to `<file>.old`, move `.new` into place.
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
