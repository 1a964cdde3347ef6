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


def test_llsafe_pm1_ecm_file_layouts(tmp_path):
    """Field orders of RunLlSafeMarin.cpp:455-477, RunPM1.cpp:6219-6223 and RunEcmTwistedEdwards.cpp:1866-1889, with this
    backend's markers; foreign or damaged files load nothing."""
    e = FakeEngine(1000)
    keep = bytes(e.blob)
    # LL-safe
    path = str(tmp_path / ck.llsafe_checkpoint_name(521))
    ck.save_llsafe(path, e, 521, 300, 275, 244, 2.5)
    raw = open(path, "rb").read()
    assert struct.unpack_from("<iIIQdQQ", raw) == (3, 521, 3, 300, 2.5, 275, 244) and raw[44:1044] == keep
    e.blob[:] = bytes(1000)
    assert ck.load_llsafe(path, e, 521) == (300, 275, 244, 2.5) and bytes(e.blob) == keep
    assert ck.load_llsafe(path, e, 523) is None
    bad = bytearray(raw); bad[8] = 1                                    # the reference's Marin tag: not this backend's image
    bad[-4:] = struct.pack("<I", (~zlib.crc32(bytes(bad[:-4])) & 0xFFFFFFFF) ^ 0xA23777AC)
    open(path, "wb").write(bad)
    assert ck.load_llsafe(path, e, 521) is None
    # P-1 stage 1
    path = str(tmp_path / ck.pm1_checkpoint_name(9941))
    st = {"checkpass": 2, "blocks_since_check": 2, "bits_in_block": 7, "current_block_len": 65, "in_lot": True, "eacc": 0xabcdef0123456789abcdef,
          "wbits": 0x55, "processed_bits": 1234, "bits": 4321}
    ck.save_pm1(path, e, 9941, 3087, st, 1.5)
    raw = open(path, "rb").read()
    assert struct.unpack_from("<iIId", raw) == (3, 9941, 3087, 1.5) and raw[20:1020] == keep
    assert struct.unpack_from("<QQQQB", raw, 1020) == (2, 2, 7, 65, 1)
    assert raw[1053:1057] == struct.pack("<I", 22) and raw[1057:1079] == b"abcdef0123456789abcdef"
    assert open(path + ".backend").read() == "marin_b200\n" and not os.path.exists(path + ".old")
    e.blob[:] = bytes(1000)
    assert ck.load_pm1(path, e, 9941) == (3087, st, 1.5) and bytes(e.blob) == keep
    open(path + ".backend", "w").write("marin\n")                       # a Marin OpenCL image under the same name
    assert ck.load_pm1(path, e, 9941) is None
    os.remove(path + ".backend")                                        # no marker = the reference's legacy Marin layout
    assert ck.load_pm1(path, e, 9941) is None
    # ECM stage 1
    path = str(tmp_path / ck.ecm_checkpoint_name(701, 4))
    ck.save_ecm(path, e, 701, 99, 8650, 6000, 0x1234567890abcdef, None, 0.25)
    raw = open(path, "rb").read()
    assert struct.unpack_from("<iIIIQdQB", raw) == (0x0301, 701, 99, 8650, 6000, 0.25, 0x1234567890abcdef, 0)
    assert ck.load_ecm(path, e, 701, 8650, 6000, 0x1234567890abcdef) == 99
    for args in ((709, 8650, 6000, 0x1234567890abcdef), (701, 8651, 6000, 0x1234567890abcdef), (701, 8650, 6001, 0x1234567890abcdef), (701, 8650, 6000, 5)):
        assert ck.load_ecm(path, e, *args) is None
    assert ck.load_ecm(path, e, 701, 8650, 6000, 0x1234567890abcdef, "16") is None


def test_mode_loops_resume_from_their_checkpoints_emulated(tmp_path):
    """LL-safe, P-1 stage 1 (Gerbicz-Li bookkeeping included) and an ECM curve, each interrupted, resumed on a NEW
    engine from the file alone and compared with the uninterrupted run."""
    from tests.emu.emu import EmuEngine
    from prmers_b200 import modes, ecm
    # LL-safe: M521 is prime; interrupted in the middle of a block
    path = str(tmp_path / ck.llsafe_checkpoint_name(521))
    r = modes.run_ll_safe(EmuEngine(521, 18), 521, check_level=2, checkpoint_path=path, stop_at=301)
    assert r["interrupted"] == 301
    log = []
    r = modes.run_ll_safe(EmuEngine(521, 18), 521, check_level=2, checkpoint_path=path, log=log.append)
    assert r["is_prime"] and any("Resuming" in s for s in log) and r["errors_gerbicz"] == 0
    whole = modes.run_ll_safe(EmuEngine(607, 18), 607, check_level=3)
    path = str(tmp_path / ck.llsafe_checkpoint_name(607))
    modes.run_ll_safe(EmuEngine(607, 18), 607, check_level=3, checkpoint_path=path, backup_every=100, stop_at=433)
    r = modes.run_ll_safe(EmuEngine(607, 18), 607, check_level=3, checkpoint_path=path, error_iter=500)   # and a fault after the resume
    assert r["res64"] == whole["res64"] and r["is_prime"] and r["errors_gerbicz"] == 1
    # P-1 stage 1 with Gerbicz-Li
    p, b1 = 2151, 256
    good = modes.run_pm1_stage1(EmuEngine(p, 11), p, b1, check_level=2)
    for stop in (37, 200, good["bits"] - 3):
        path = str(tmp_path / f"s{stop}_{ck.pm1_checkpoint_name(p)}")
        r = modes.run_pm1_stage1(EmuEngine(p, 11), p, b1, check_level=2, checkpoint_path=path, stop_at=stop)
        assert r["interrupted"] == stop
        log = []
        r = modes.run_pm1_stage1(EmuEngine(p, 11), p, b1, check_level=2, checkpoint_path=path, log=log.append, error_iter=stop + 2)
        assert r["x"] == good["x"] and r["factor"] == 327405968242246366421788399 and r["errors_gerbicz"] == 1
        assert any("Resuming" in s for s in log)
    # ECM curve
    p, seed = 101, ecm.mix64(123456789, 0)
    whole = ecm.run_curve(EmuEngine(p, ecm.REGS), p, 300, 0, seed)
    path = str(tmp_path / ck.ecm_checkpoint_name(p, 0))
    r = ecm.run_curve(EmuEngine(p, ecm.REGS), p, 300, 0, seed, checkpoint_path=path, backup_every=50, stop_at=123)
    assert r["status"] == "interrupted" and r["step"] == 123
    log = []
    r = ecm.run_curve(EmuEngine(p, ecm.REGS), p, 300, 0, seed, checkpoint_path=path, log=log.append)
    assert (r["status"], r["factor"], r["stage1_T"]) == (whole["status"], whole["factor"], whole["stage1_T"]) and any("Resuming" in s for s in log)
    assert ecm.run_curve(EmuEngine(p, ecm.REGS), p, 300, 0, seed + 1, checkpoint_path=path, log=log.append)["status"] in ("NF", "found")   # another curve ignores the file


def test_ecm_stage2_checkpoints_emulated(tmp_path):
    """Stage 2 of a curve interrupted inside a chunk, inside a later chunk and at a chunk boundary resumes from the file
    alone (fresh engine) to the result of the uninterrupted run; a finished stage 2 leaves no file; a file of another
    curve is ignored (RunEcmTwistedEdwards.cpp:1940-2078, 3640-3664, 3797-3822)."""
    from tests.emu.emu import EmuEngine
    from prmers_b200 import ecm
    p, b1, b2, bits = 701, 50, 400, 300
    for seed_idx, expect in ((1, "NF"), (3, "found")):
        seed = ecm.mix64(7, seed_idx)
        whole = ecm.run_curve(EmuEngine(p, ecm.REGS), p, b1, b2, seed, s2_bits=bits)
        assert whole["status"] == expect and (expect == "NF" or whole["stage"] == 2)
        path = str(tmp_path / ck.ecm2_checkpoint_name(p, seed_idx))
        for stop in ((0, 40), (1, 10), (0, None)):
            if expect == "found" and stop[0] == 1 and whole["point_ops"] < 600:
                continue                                     # the factor appears at the end of the first chunk
            r = ecm.run_curve(EmuEngine(p, ecm.REGS), p, b1, b2, seed, s2_bits=bits, checkpoint2_path=path, backup_every=25, stop_at2=stop)
            if r["status"] != "interrupted":
                assert expect == "found" and (r["status"], r["factor"]) == (whole["status"], whole["factor"])
                continue
            assert r["stage"] == 2 and os.path.exists(path)
            log = []
            assert ecm.run_curve(EmuEngine(p, ecm.REGS), p, b1, b2, seed + 2, s2_bits=bits, checkpoint2_path=path + ".other")["status"] in ("NF", "found")
            r = ecm.run_curve(EmuEngine(p, ecm.REGS), p, b1, b2, seed, s2_bits=bits, checkpoint2_path=path, log=log.append)
            assert (r["status"], r["factor"], r.get("stage2_T")) == (whole["status"], whole["factor"], whole.get("stage2_T")), stop
            assert any("Resuming stage 2" in s for s in log) and not os.path.exists(path)
    # a file of another curve / other bounds is not loaded
    seed = ecm.mix64(7, 1)
    path = str(tmp_path / "foreign.ckpt")
    assert ecm.run_curve(EmuEngine(p, ecm.REGS), p, b1, b2, seed, s2_bits=bits, checkpoint2_path=path, stop_at2=(0, 30))["status"] == "interrupted"
    log = []
    r = ecm.run_curve(EmuEngine(p, ecm.REGS), p, b1, b2 + 100, seed, s2_bits=bits, checkpoint2_path=path, log=log.append)
    assert not any("Resuming stage 2" in s for s in log) and r["status"] in ("NF", "found")


def test_ecm_curve_loop_with_checkpoint_files_emulated(tmp_path):
    """run_ecm with a checkpoint directory: interrupted in stage 1 of one curve, then in stage 2 of another, the same call
    made again carries on from the files and ends with the results of the uninterrupted loop."""
    from tests.emu.emu import EmuEngine
    from prmers_b200 import ecm
    p, b1, b2, bits, curves, seed = 701, 50, 400, 300, 4, 7
    fac = lambda q, regs: EmuEngine(q, regs)
    key = lambda out: [(c["curve"], c["status"], c["factor"], c.get("stage2_T")) for c in out["curves"]]
    whole = ecm.run_ecm(fac, p, b1, b2, curves=curves, seed=seed, continue_after_factor=True, s2_bits=bits)
    d = str(tmp_path)
    done = []
    for stop in ((0, 1, 20), (1, 2, (0, 35)), (2, 2, (1, 5)), None):
        out = ecm.run_ecm(fac, p, b1, b2, curves=curves, seed=seed, continue_after_factor=True, s2_bits=bits, checkpoint_dir=d, backup_every=10,
                          stop=stop, curve_indices=range(len(done), curves))
        if stop is not None and out["status"] == "interrupted":
            assert out["interrupted_curve"] == stop[0]
            done += [c for c in out["curves"] if c["status"] != "interrupted"]
        else:
            done += out["curves"]
            if stop is None:
                break
    assert [(c["curve"], c["status"], c["factor"], c.get("stage2_T")) for c in done] == key(whole)
    assert not [f for f in os.listdir(d) if f.endswith(".ckpt")]


@pytest.mark.gpu
def test_ecm_stage1_and_stage2_checkpoints_gpu(tmp_path):
    """The same on the device (recorded point operations, i.e. programs with renamed registers: a checkpoint between two
    replays holds everything that is live): stage 1 interrupted and resumed, then stage 2 interrupted inside its second
    chunk and resumed, equals the uninterrupted run."""
    import prmers_b200 as pb
    from prmers_b200 import ecm
    p, b1, b2, bits = 701, 50, 400, 300
    seed = ecm.mix64(7, 1)
    make = lambda: pb.Engine(p, ecm.REGS)
    whole = ecm.run_curve(make(), p, b1, b2, seed, s2_bits=bits)
    p1, p2 = str(tmp_path / ck.ecm_checkpoint_name(p, 1)), str(tmp_path / ck.ecm2_checkpoint_name(p, 1))
    r = ecm.run_curve(make(), p, b1, b2, seed, s2_bits=bits, checkpoint_path=p1, checkpoint2_path=p2, stop_at=33)
    assert r["status"] == "interrupted" and r["stage"] == 1
    r = ecm.run_curve(make(), p, b1, b2, seed, s2_bits=bits, checkpoint_path=p1, checkpoint2_path=p2, stop_at2=(1, 10))
    assert r["status"] == "interrupted" and r["stage"] == 2 and r["chunk"] == 1
    log = []
    r = ecm.run_curve(make(), p, b1, b2, seed, s2_bits=bits, checkpoint_path=p1, checkpoint2_path=p2, log=log.append)
    assert (r["status"], r["factor"], r["stage2_T"]) == (whole["status"], whole["factor"], whole["stage2_T"])
    assert any("Resuming stage 2" in s for s in log) and not os.path.exists(p2)
