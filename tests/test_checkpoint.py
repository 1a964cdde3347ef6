"""Checkpoint container (prmers_b200/checkpoint.py): format, CRC rule of include/marin/file.h, rotation,
rejection of foreign / damaged files; and on the GPU a PRP run interrupted and resumed from the file."""
import os
import struct
import zlib

import pytest

from prmers_b200 import checkpoint as ck


class FakeEngine:
    def __init__(self, size=4096):
        self.blob = bytearray(os.urandom(size))

    def get_checkpoint_size(self):
        return len(self.blob)

    def get_checkpoint(self, data):
        if len(data) != len(self.blob):
            return False
        data[:] = self.blob
        return True

    def set_checkpoint(self, data):
        if len(data) != len(self.blob):
            return False
        self.blob[:] = data
        return True


def test_format_crc_and_rotation(tmp_path):
    path = str(tmp_path / ck.checkpoint_name(9941))
    e = FakeEngine()
    ck.save(path, e, 9941, 1234, 5.5)
    raw = open(path, "rb").read()
    version, p, mode, backend, it, et = struct.unpack_from("<iIIIId", raw)
    assert (version, p, mode, backend, it, et) == (2, 9941, 1, 3, 1234, 5.5)
    # file.h:96-110: stored word = ~crc32(everything before) ^ 0xa23777ac
    assert struct.unpack_from("<I", raw, len(raw) - 4)[0] == (~zlib.crc32(raw[:-4]) & 0xFFFFFFFF) ^ 0xA23777AC
    first = bytes(e.blob)
    e.blob[:] = os.urandom(len(e.blob))
    ck.save(path, e, 9941, 2000, 7.0)
    assert os.path.exists(path + ".old") and not os.path.exists(path + ".new")
    second = bytes(e.blob)
    e.blob[:] = bytes(len(e.blob))
    assert ck.load_latest(path, e, 9941) == (2000, 7.0) and bytes(e.blob) == second
    os.remove(path)
    assert ck.load_latest(path, e, 9941) == (1234, 5.5) and bytes(e.blob) == first


def test_foreign_and_damaged_files_are_ignored(tmp_path):
    path = str(tmp_path / "m_9941.ckpt")
    e = FakeEngine()
    ck.save(path, e, 9941, 10)
    keep = bytes(e.blob)
    assert ck.load(path, e, 9949) is None                        # other exponent
    assert ck.load(path, e, 9941, mode=ck.MODE_LL) is None       # other mode
    raw = bytearray(open(path, "rb").read())
    raw[12] = 1                                                  # backend tag 1 = Marin OpenCL image
    open(path, "wb").write(raw)
    assert ck.load(path, e, 9941) is None
    raw[12] = 3; raw[100] ^= 0xFF                                # bit rot
    open(path, "wb").write(raw)
    assert ck.load(path, e, 9941) is None
    assert bytes(e.blob) == keep                                 # nothing was loaded
    assert ck.load(str(tmp_path / "missing.ckpt"), e, 9941) is None


@pytest.mark.gpu
def test_prp_interrupt_and_resume_gpu(tmp_path):
    import prmers_b200 as pb
    p = 9941
    path = str(tmp_path / ck.checkpoint_name(p))
    e = pb.Engine(p, 8)
    e.set(0, 3); e.set(1, 1)
    e.square_mul_repeat(0, 5000)
    ck.save(path, e, p, 5000, 1.0)
    e.close()
    e2 = pb.Engine(p, 8)
    assert ck.load_latest(path, e2, p) == (5000, 1.0)
    e2.square_mul_repeat(0, p - 5000)
    assert e2.digit(0).equal_to(9)                               # M9941 is prime
    e2.close()
