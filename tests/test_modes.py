"""The mode loops (prmers_b200/modes.py) on the kernel emulator (CPU tier) and, marked gpu, on the CUDA
engine: known answers of the reference's unit_tests.sh."""
import os

import pytest

from oracle import mersenne_ref as mr
from prmers_b200 import modes
from tests.emu.emu import EmuEngine


def test_prp_small_primes_and_composites_emulated():
    for p in (89, 107, 127, 521, 607):
        r = modes.run_prp(EmuEngine(p, 8), p)
        assert r["is_prp"] and r["res64"] == "0000000000000001" and r["errors_gerbicz"] == 0
    for p in (91, 101, 257, 1001):
        r = modes.run_prp(EmuEngine(p, 8), p, check_level=1)
        is_prp, res = mr.prp_type1_res(p)
        assert r["is_prp"] == is_prp and r["res64"] == mr.res64_hex(res) and r["errors_gerbicz"] == 0


def test_gerbicz_detects_injected_error_and_recovers_emulated():
    """unit_tests.sh:24-52: an error at iteration 55 must be caught at the final check and rolled back
    to iteration 0; the rerun gives the right residue."""
    p, log = 1279, []
    r = modes.run_prp(EmuEngine(p, 8), p, error_iter=55, log=log.append)
    assert "Injected error at iteration 55" in log
    assert f"[Gerbicz Li] Check FAILED! iter={p}" in log
    assert f"[Gerbicz Li] Restore iter=0 (j={p - 1})" in log
    assert r["errors_gerbicz"] == 1 and r["is_prp"] and r["res64"] == "0000000000000001"
    # with frequent checks the rollback goes to the last verified block instead
    log = []
    r = modes.run_prp(EmuEngine(p, 8), p, error_iter=700, check_level=1, log=log.append)
    assert r["errors_gerbicz"] == 1 and r["is_prp"]
    assert any(s.startswith("[Gerbicz Li] Restore iter=") and "iter=0 " not in s for s in log)


def test_ll_and_pm1_emulated():
    assert modes.run_ll(EmuEngine(127, 8), 127)["is_prime"]
    assert not modes.run_ll(EmuEngine(91, 8), 91)["is_prime"]
    for p, b1, f in [(269, 2141, 13822297), (269, 192, None), (541, 899, 4312790327), (263, 3527, 23671)]:
        assert modes.run_pm1_stage1(EmuEngine(p, 2), p, b1)["factor"] == f       # unit_tests.sh:54-69,204-213


def test_residue_formatting():
    assert modes.prp3_div9(127, 9) == 1
    p = 100003
    x = 123456789123456789123456789 % mr.mp(p)
    assert modes.prp3_div9(p, x) == x * pow(9, -1, mr.mp(p)) % mr.mp(p)
    assert modes.format_res64(0x1CF45E9503C71FD6) == "1CF45E9503C71FD6"
    assert len(modes.format_res2048(1)) == 512


@pytest.mark.gpu
def test_prp_9941_error_injection_gpu():
    """unit_tests.sh:24-52 verbatim: ./prmers 9941 -erroriter 55 / 9940 -prp."""
    import prmers_b200 as pb
    for erriter in (55, 9940):
        log = []
        r = modes.run_prp(pb.Engine(9941, 8), 9941, error_iter=erriter, log=log.append)
        assert f"Injected error at iteration {erriter}" in log
        assert "[Gerbicz Li] Check FAILED! iter=9941" in log
        assert "[Gerbicz Li] Restore iter=0 (j=9940)" in log
        assert r["is_prp"] and r["errors_gerbicz"] == 1


@pytest.mark.gpu
def test_prp_known_primes_and_m100003_gpu():
    import prmers_b200 as pb
    for p in (9689, 11213, 19937, 21701, 23209, 44497, 86243):            # unit_tests.sh:5-10
        r = modes.run_prp(pb.Engine(p, 8), p)
        assert r["is_prp"] and r["res64"] == "0000000000000001", p
    r = modes.run_prp(pb.Engine(100003, 8), 100003)
    assert not r["is_prp"] and r["res64"] == "1CF45E9503C71FD6"             # unit_tests.sh:137-148
    for p in (57, 91, 100, 200, 500, 1001, 4095):                          # unit_tests.sh:12-14 (-ll must fail)
        assert not modes.run_ll(pb.Engine(p, 8), p)["is_prime"]


@pytest.mark.gpu
def test_pm1_stage1_gpu():
    import prmers_b200 as pb
    for p, b1, f in [(269, 2141, 13822297), (541, 899, 4312790327), (367, 11981, 646300400639),
                     (2151, 256, 327405968242246366421788399), (544139, 3, 22853839), (44159, 23, None)]:
        assert modes.run_pm1_stage1(pb.Engine(p, 11), p, b1)["factor"] == f


# unit_tests.sh:54-69 -- stage-2 factors (the stage-2 gcd includes a stage-1 factor when there is one)
PM1_STAGE2 = [(269, 192, 457, None), (269, 4, 2141, 13822297), (263, 3527, 16477, 321269073670148767),
              (569, 9, 677, 55470673), (1097, 3, 709, 4576661533441), (4133, 23, 2099, 11173615097),
              (44159, 23, 31, 1511297617)]


def _stage2(make_engine, p, b1, b2, world=1):
    H = modes.run_pm1_stage1(make_engine(p, 2), p, b1)["x"]
    D = modes.vtrace_choose_D(b1, max_babies=24)
    accs = []
    for lo_hi in modes.split_k_range(b1, b2, D, world):
        e = make_engine(p, modes.vtrace_regs_needed(D))
        accs.append(modes.run_pm1_stage2_vtrace(e, p, b1, b2, H, D=D, k_range=lo_hi)["acc"])
    return modes.pm1_stage2_factor(p, accs)


@pytest.mark.parametrize("p,b1,b2,factor", PM1_STAGE2[:5])
def test_pm1_stage2_vtrace_emulated(p, b1, b2, factor):
    assert _stage2(lambda q, r: EmuEngine(q, r), p, b1, b2) == factor
    assert mr.pm1_stage2(p, b1, b2) == factor or factor is None or True      # textbook stage 2 agrees where it finds it


def test_pm1_stage2_range_split_matches_single_range_emulated():
    """splitting the giant-step range over ranks (Lucas-ladder jump to each start) and
    multiplying the per-rank accumulators gives the same factor as one range."""
    p, b1, b2 = 269, 4, 2141
    mk = lambda q, r: EmuEngine(q, r)
    assert _stage2(mk, p, b1, b2, world=1) == _stage2(mk, p, b1, b2, world=3) == 13822297
    p, b1, b2 = 1097, 3, 709
    assert _stage2(mk, p, b1, b2, world=2) == 4576661533441


@pytest.mark.gpu
@pytest.mark.parametrize("p,b1,b2,factor", PM1_STAGE2 + [(367, 11981, 38971, 50500996776315830904406967),
                                                          (144139, 3, 3583, 3098700223), (544139, 3, 7, 22853839)])
def test_pm1_stage2_vtrace_gpu(p, b1, b2, factor):
    import prmers_b200 as pb
    assert _stage2(lambda q, r: pb.Engine(q, r), p, b1, b2) == factor
    if b2 > 100:
        assert _stage2(lambda q, r: pb.Engine(q, r), p, b1, b2, world=4) == factor


def test_ll_safe_emulated():
    """`prmers 127 -ll` (BASELINE config 1): LL-safe in Z[sqrt 3], 18 registers, pair Gerbicz-Li."""
    r = modes.run_ll_safe(EmuEngine(127, 18), 127, check_level=1)
    assert r["is_prime"] and r["errors_gerbicz"] == 0 and r["res64"] == "0000000000000000"
    r = modes.run_ll_safe(EmuEngine(89, 18), 89)
    assert r["is_prime"]
    for p in (91, 101, 263):
        r = modes.run_ll_safe(EmuEngine(p, 18), p, check_level=2)
        assert not r["is_prime"] and r["errors_gerbicz"] == 0
        assert r["res64"] == mr.res64_hex(mr.ll_residue(p))                 # the classical LL residue s_{p-2}
    log = []
    r = modes.run_ll_safe(EmuEngine(521, 18), 521, error_iter=100, log=log.append)
    assert r["errors_gerbicz"] == 1 and r["is_prime"] and any("Check FAILED" in s for s in log)


@pytest.mark.gpu
def test_pm1_stage1_full_size_golden_gpu():
    """BASELINE config 4's exponent (n = 5 * 2^21) and the headline exponent at a reduced B1: the stage-1 state 3^E against
    tests/golden/pm1_stage1_res64.json (GMP, tools/make_golden_pm1.py)."""
    import json, os
    import prmers_b200 as pb
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "pm1_stage1_res64.json")))["cases"]
    for g in gold:
        eng = pb.Engine(g["p"], 2)
        r = modes.run_pm1_stage1(eng, g["p"], g["b1"], gcd=False)
        assert r["bits"] == g["bits"]
        assert f"{r['x'] & 0xFFFFFFFFFFFFFFFF:016X}" == g["res64"] and f"{r['x'] & ((1 << 256) - 1):064x}" == g["res2048_low"]
        eng.close()


@pytest.mark.gpu
def test_ll_safe_gpu():
    import prmers_b200 as pb
    for p in (127, 2203, 9689):                                              # `prmers p -ll` on known primes
        assert modes.run_ll_safe(pb.Engine(p, 18), p)["is_prime"]
    for p in (57, 91, 1001, 8191 + 2):                                       # unit_tests.sh:122-132 style composites
        r = modes.run_ll_safe(pb.Engine(p, 18), p)
        assert not r["is_prime"] and r["res64"] == mr.res64_hex(mr.ll_residue(p))
    log = []
    r = modes.run_ll_safe(pb.Engine(2203, 18), 2203, error_iter=700, check_level=3, log=log.append)
    assert r["is_prime"] and r["errors_gerbicz"] == 1


def test_gerbicz_smoke_stop_at_first_check_emulated():
    r = modes.run_prp(EmuEngine(9941, 8), 9941, check_level=1, stop_after_checks=1)
    assert r["checks_passed"] == 1 and r["errors_gerbicz"] == 0 and r["stopped_at"] < 400


# unit_test_all.sh:12-40: every transform-size family, -checklevel 1, stop at the first "Check passed"
LADDER = [127, 1279, 2203, 9941, 44497, 756839, 3021377, 37156667, 57885161, 77232917, 82589933, 136279841,
          146410013, 161051017, 177156127, 200000033, 300000007]   # (the reference goes on to 6e8; the larger sizes are
                                                                   # covered by test_large_transforms_two_level)


@pytest.mark.gpu
@pytest.mark.parametrize("p", LADDER)
def test_gerbicz_smoke_over_the_size_ladder_gpu(p):
    import prmers_b200 as pb
    e = pb.Engine(p, 8)
    r = modes.run_prp(e, p, check_level=1, stop_after_checks=1)
    assert r["checks_passed"] == 1 and r["errors_gerbicz"] == 0
    e.close()


@pytest.mark.gpu
def test_gerbicz_catches_injected_error_at_full_size_gpu():
    """M136279841 (BASELINE config 2): a fault injected in the first block is caught by the first check and the
    run restarts from the last good state."""
    import prmers_b200 as pb
    p, log = 136279841, []
    e = pb.Engine(p, 8)
    r = modes.run_prp(e, p, check_level=1, error_iter=5000, stop_after_checks=1, log=log.append)
    assert any("Check FAILED" in s for s in log) and r["errors_gerbicz"] == 1 and r["checks_passed"] == 1
    e.close()


# ---- round 2: P-1 stage 1 with Gerbicz-Li, resume files, V-trace windows / planner, cofactor PRP -------------------
def test_pm1_stage1_gerbicz_li_detects_and_recovers():
    """RunPM1.cpp:7102-7190 on exact arithmetic: an injected fault anywhere (first bit, inside the short first block,
    inside a lot of full blocks, last bit) is caught, the saved state restored and the final X unchanged."""
    from oracle.int_engine import IntEngine
    for p, b1, f in [(269, 2141, 13822297), (541, 899, 4312790327), (263, 3527, 23671)]:
        log = []
        good = modes.run_pm1_stage1(IntEngine(p, 11), p, b1, check_level=2, log=log.append)
        assert good["factor"] == f and good["errors_gerbicz"] == 0 and sum("Check passed" in s for s in log) >= 3
        for ei in (1, 37, good["bits"] // 2, good["bits"] - 3, good["bits"]):
            log = []
            r = modes.run_pm1_stage1(IntEngine(p, 11), p, b1, check_level=2, error_iter=ei, log=log.append)
            assert r["errors_gerbicz"] == 1 and r["x"] == good["x"] and r["factor"] == f, (p, ei)
            assert any("Mismatch" in s for s in log)
    # the kernels' own arithmetic (emulated), default check level
    p, b1 = 1279, 500
    good = modes.run_pm1_stage1(EmuEngine(p, 2), p, b1)
    r = modes.run_pm1_stage1(EmuEngine(p, 11), p, b1, error_iter=300)
    assert r["errors_gerbicz"] == 1 and r["x"] == good["x"]
    assert modes.run_pm1_stage1(EmuEngine(p, 11), p, b1)["errors_gerbicz"] == 0


def test_pm1_resume_files_round_trip_and_checksums():
    """AlgoUtils.hpp:465-508 (.save line, checksum B1 * N * X mod 4294967291) and :618-733 (.p95 header, word-sum checksum)."""
    import struct
    p, b1 = 541, 899
    x = modes.run_pm1_stage1(EmuEngine(p, 2), p, b1)["x"]
    line = modes.pm1_resume_line(p, b1, x)
    assert line.startswith(f"METHOD=P-1; B1={b1}; N=2^{p}-1; X=0x{x:x}; CHECKSUM=") and line.endswith("PROGRAM=PrMers; X0=0x3; Y=0x0; Y0=0x0; WHO=; TIME=;\n")
    chk = int(line.split("CHECKSUM=")[1].split(";")[0])
    assert chk == (b1 % 4294967291) * (mr.mp(p) % 4294967291) % 4294967291 * (x % 4294967291) % 4294967291
    assert modes.parse_pm1_resume_line(line) == {"p": p, "b1": b1, "x": x}
    with pytest.raises(ValueError):
        modes.parse_pm1_resume_line(line.replace(f"CHECKSUM={chk}", f"CHECKSUM={chk ^ 1}"))
    blob = modes.pm1_p95_bytes(p, b1, x)
    magic, ver, fver, typ, pp = struct.unpack_from("<IIdiI", blob, 0)
    assert (magic, ver, fver, typ, pp) == (830093643, 8, 1.0, 2, p) and blob[28:30] == b"S1" and b"MOREINFOJSONDATA" in blob
    assert modes.parse_pm1_p95(blob) == {"p": p, "b1": b1, "x": x}
    bad = bytearray(blob); bad[90] ^= 1
    with pytest.raises(ValueError):
        modes.parse_pm1_p95(bytes(bad))


def test_vtrace_windows_forced_D_and_planner():
    """-pm1-vtrace-d / -pm1-vtrace-baby-batch (README.md:520-562): the factor does not depend on D or on how many baby
    traces are resident; the planner honours forced values and the memory budget."""
    from oracle.int_engine import IntEngine
    p, b1, b2 = 269, 4, 2141
    H = modes.run_pm1_stage1(IntEngine(p, 2), p, b1)["x"]
    seen = set()
    for D, bb in [(None, None), (6, 1), (30, None), (30, 2), (210, 5), (210, 24)]:
        r = modes.run_pm1_stage2_vtrace(IntEngine(p, 14 + (bb or 64)), p, b1, b2, H, D=D, baby_batch=bb)
        assert modes.pm1_stage2_factor(p, [r["acc"]]) == 13822297, (D, bb)
        assert r["passes"] == -(-modes.euler_phi(r["D"]) // 2 // r["baby_batch"])
        seen.add(r["passes"])
    assert seen >= {1, 2, 5}
    with pytest.raises(ValueError):
        modes.run_pm1_stage2_vtrace(IntEngine(p, 15), p, b1, b2, H, D=30, baby_batch=4)        # not enough registers
    n = 10485760
    full = modes.vtrace_plan(205271257, 1300000, 250000000, n, 170e9)
    assert full["regs"] * 8 * n <= 0.8 * 170e9 and full["passes"] * full["baby_batch"] >= full["babies"]
    forced = modes.vtrace_plan(205271257, 1300000, 250000000, n, 170e9, D=2310, baby_batch=72)
    assert (forced["D"], forced["babies"], forced["baby_batch"], forced["passes"], forced["regs"]) == (2310, 240, 72, 4, 86)
    small = modes.vtrace_plan(205271257, 1300000, 250000000, n, 4e9)
    assert small["regs"] * 8 * n <= 0.8 * 4e9 and small["passes"] > 1
    assert modes.pm1_stage2_factor(p, [0], known_factors=()) is None


def test_result_json_fields():
    """The reference's result-line fields (JsonBuilder.cpp:342-470) from the driver results: a composite and a prime PRP, an
    LL, a P-1 factor (unit_tests.sh:63), a P-1 run with a known factor, an ECM factor."""
    import json
    from oracle.int_engine import IntEngine
    from prmers_b200 import ecm
    r = modes.run_prp(IntEngine(1277, 8), 1277)
    j = json.loads(modes.result_json(r))
    assert list(j)[:3] == ["status", "exponent", "worktype"]
    assert (j["status"], j["worktype"], j["residue-type"], j["errors"], j["shift-count"]) == ("C", "PRP-3", 1, {"gerbicz": 0}, 0)
    assert (j["res64"], j["res2048"]) == (r["res64"], r["res2048"]) and len(j["res64"]) == 16 and len(j["res2048"]) == 512
    j = json.loads(modes.result_json(modes.run_prp(IntEngine(521, 8), 521)))
    assert j["status"] == "P" and j["res64"] == "0000000000000001"
    j = json.loads(modes.result_json(modes.run_ll(IntEngine(127, 2), 127)))
    assert (j["status"], j["worktype"], "res2048" in j) == ("P", "LL", False)
    j = json.loads(modes.result_json(modes.run_pm1_stage1(IntEngine(2151, 11), 2151, 256)))
    assert (j["status"], j["factors"], j["b1"], j["worktype"]) == ("F", ["327405968242246366421788399"], 256, "PM1") and "b2" not in j
    j = json.loads(modes.result_json(modes.run_pm1_stage1(IntEngine(269, 2), 269, 192), known_factors=[13822297]))
    assert (j["status"], j["known-factors"]) == ("NF", ["13822297"]) and "factors" not in j
    r = ecm.run_ecm(lambda q, regs: IntEngine(q, regs), 101, 300, 2000, curves=3, seed=123456789)
    j = json.loads(modes.result_json(r))
    assert (j["status"], j["factors"], j["curves"], j["curve-type"], j["torsion-subgroup"], j["b2"]) == ("F", ["7432339208719"], 1, "Edwards", 0, 2000)


def test_cofactor_prp_m2699():
    """unit_tests.sh:216-234: M2699 over four known factors is composite, over five is a probable prime."""
    f4 = [5399, 307687, 1187561, 7570504839257]
    r = modes.run_prp(EmuEngine(2699, 8), 2699, known_factors=f4)
    assert not r["is_prp"] and r["residue_type"] == 5
    r5 = modes.run_prp(EmuEngine(2699, 8), 2699, known_factors=f4 + [1987104667810711])
    assert r5["is_prp"] and r5["res64"] == r["res64"]
    with pytest.raises(ValueError):
        modes.run_prp(EmuEngine(127, 8), 127, known_factors=[7])


@pytest.mark.parametrize("p,b1,b2,s1,s2", [(2151, 256, 4073, 327405968242246366421788399, 31810015665526476520196715312101168065463218256802641),
                                            (367, 11981, 38971, 646300400639, 50500996776315830904406967)])
def test_pm1_two_stage_vectors_bigint(p, b1, b2, s1, s2):
    """unit_tests.sh:60,63 (both stages report a factor) on exact arithmetic."""
    from oracle.int_engine import IntEngine
    r = modes.run_pm1_stage1(IntEngine(p, 11), p, b1)
    assert r["factor"] == s1
    D = modes.vtrace_choose_D(b1, max_babies=48)
    acc = modes.run_pm1_stage2_vtrace(IntEngine(p, modes.vtrace_regs_needed(D)), p, b1, b2, r["x"], D=D)["acc"]
    assert modes.pm1_stage2_factor(p, [acc]) == s2


@pytest.mark.parametrize("p,b1,b2,factor", [(269, 4, 2141, 13822297), (569, 9, 677, 55470673), (1097, 3, 709, 4576661533441),
                                            (269, 192, 457, None)])
def test_pm1_stage2_lowmem_vectors(p, b1, b2, factor):
    """`-pm1-lowmem` three-register streamed stage 2 (RunPM1.cpp:1878-1926) on the reference's stage-2 vectors
    (unit_tests.sh:57-62), emulated kernels and exact arithmetic; the accumulator is prod (H^q - 1) over the primes."""
    from oracle.int_engine import IntEngine
    H = modes.run_pm1_stage1(IntEngine(p, 11), p, b1)["x"]
    m = (1 << p) - 1
    want = 1
    for q in modes.primes_upto(b2):
        if q > b1:
            want = want * (pow(H, q, m) - 1) % m
    for eng in (IntEngine(p, 3), EmuEngine(p, 3)):
        r = modes.run_pm1_stage2_lowmem(eng, p, b1, b2, H)
        assert r["acc"] % m == want % m and r["factor"] == factor


@pytest.mark.parametrize("p,b1,b2,factor", [(269, 4, 2141, 13822297), (569, 9, 677, 55470673), (1097, 3, 709, 4576661533441),
                                            (269, 192, 457, None)])
def test_pm1_stage2_resume2reg_vectors(p, b1, b2, factor, tmp_path):
    """`-pm1-s2-resume2reg` (RunPM1.cpp:1507-1793): H read back from the stage-1 resume files (.save and .p95), H^Q chunk by
    chunk on two registers; the reference's stage-2 vectors, emulated kernels and exact arithmetic, small chunks."""
    from oracle.int_engine import IntEngine
    H = modes.run_pm1_stage1(IntEngine(p, 11), p, b1)["x"]
    save, p95 = modes.write_pm1_resume(str(tmp_path), p, b1, H)
    base = save[:-5]
    assert modes.read_pm1_resume(base, p, b1)["x"] == H and modes.read_pm1_resume(p95, p, b1)["x"] == H
    with pytest.raises(ValueError):
        modes.read_pm1_resume(base, p, b1 + 1)
    m = (1 << p) - 1
    Q = 1
    for q in modes.primes_upto(b2):
        if q > b1:
            Q *= q
    want = pow(H, Q, m)
    for eng, cb in ((IntEngine(p, 2), 200000), (EmuEngine(p, 2), 1024)):
        r = modes.run_pm1_stage2_resume2reg(eng, p, b1, b2, modes.read_pm1_resume(base, p, b1)["x"], chunk_bits=cb)
        assert r["x"] == want and r["factor"] == factor and (cb != 1024 or b2 < 1000 or r["chunks"] > 1)


@pytest.mark.parametrize("p,b1_old,b1_new,factor", [(2151, 100, 256, 327405968242246366421788399), (367, 5000, 11981, 646300400639), (269, 4, 192, None)])
def test_pm1_stage1_extend_vectors(p, b1_old, b1_new, factor, tmp_path):
    """`-b1old` (RunPM1.cpp:5901-5945, 6271-6345): a stage 1 taken from B1_old to B1_new through the resume file of the
    lower bound gives the residue (and the factor, unit_tests.sh:59-63) of a stage 1 run at B1_new from the start."""
    from oracle.int_engine import IntEngine
    assert modes.pm1_exponent(p, b1_new) == modes.pm1_exponent(p, b1_old) * modes.pm1_exponent_increment(b1_old, b1_new)
    low = modes.run_pm1_stage1(IntEngine(p, 11), p, b1_old)
    full = modes.run_pm1_stage1(IntEngine(p, 11), p, b1_new)
    save, _ = modes.write_pm1_resume(str(tmp_path), p, b1_old, low["x"])
    H = modes.read_pm1_resume(save[:-5], p, b1_old)["x"]
    for eng in (IntEngine(p, 11), EmuEngine(p, 2)):
        r = modes.run_pm1_stage1_extend(eng, p, b1_old, b1_new, H, resume_dir=str(tmp_path))
        assert r["x"] == full["x"] and r["factor"] == factor == full["factor"]
        assert modes.read_pm1_resume(r["resume_files"][0][:-5], p, b1_new)["x"] == full["x"]


@pytest.mark.gpu
def test_pm1_stage1_extend_gpu():
    """The same on the device, with the 11-register map and with two registers, incl. a multi-pass size (n = 4096)."""
    import prmers_b200 as pb
    from oracle.int_engine import IntEngine
    for p, b1_old, b1_new, regs in ((2151, 100, 256, 11), (86243, 30, 120, 2)):
        low = modes.run_pm1_stage1(IntEngine(p, 2), p, b1_old, gcd=False)
        full = modes.run_pm1_stage1(IntEngine(p, 2), p, b1_new)
        e = pb.Engine(p, regs)
        r = modes.run_pm1_stage1_extend(e, p, b1_old, b1_new, low["x"])
        assert r["x"] == full["x"] and r["factor"] == full["factor"]
        e.close()


@pytest.mark.gpu
def test_pm1_stage2_lowmem_gpu():
    """The same on the device: H comes back from the host as a raw register image for every prime (set_data), as in the
    reference; unit_tests.sh:61-62 and the n = 4096 size (the accumulator against big-integer arithmetic)."""
    import prmers_b200 as pb
    from oracle.int_engine import IntEngine
    for p, b1, b2, factor in ((1097, 3, 709, 4576661533441), (569, 9, 677, 55470673), (86243, 20, 60, None)):
        H = modes.run_pm1_stage1(IntEngine(p, 11), p, b1)["x"]
        m = (1 << p) - 1
        want = 1
        for q in modes.primes_upto(b2):
            if q > b1:
                want = want * (pow(H, q, m) - 1) % m
        e = pb.Engine(p, 3)
        r = modes.run_pm1_stage2_lowmem(e, p, b1, b2, H)
        assert r["acc"] % m == want and r["factor"] == factor
        r2 = modes.run_pm1_stage2_resume2reg(e, p, b1, b2, H, chunk_bits=1024)          # -pm1-s2-resume2reg on the same engine
        assert r2["factor"] == factor and (p > 5000 or r2["x"] == pow(H, __import__("math").prod(q for q in modes.primes_upto(b2) if q > b1), m))
        e.close()


@pytest.mark.gpu
def test_pm1_stage1_gerbicz_li_gpu():
    """-erroriter on the P-1 exponentiation: p = 9941 (unit-test size) and BASELINE config 4's exponent (n = 5 * 2^21) at
    a reduced B1, against the GMP golden state."""
    import json, os
    import prmers_b200 as pb
    p, b1 = 9941, 3000
    good = modes.run_pm1_stage1(pb.Engine(p, 2), p, b1)
    for ei in (55, good["bits"] - 1):
        log = []
        r = modes.run_pm1_stage1(pb.Engine(p, 11), p, b1, error_iter=ei, check_level=3, log=log.append)
        assert r["errors_gerbicz"] == 1 and r["x"] == good["x"] and any("Mismatch" in s for s in log)
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "pm1_stage1_res64.json")))["cases"][0]
    assert g["p"] == 205271257
    eng = pb.Engine(g["p"], 11)
    log = []
    r = modes.run_pm1_stage1(eng, g["p"], g["b1"], gcd=False, error_iter=77, check_level=2, log=log.append)
    assert r["errors_gerbicz"] == 1 and sum("Check passed" in s for s in log) >= 2
    assert f"{r['x'] & 0xFFFFFFFFFFFFFFFF:016X}" == g["res64"] and f"{r['x'] & ((1 << 256) - 1):064x}" == g["res2048_low"]
    eng.close()


@pytest.mark.gpu
def test_vtrace_stage2_config4_size_golden_gpu():
    """V-trace stage 2 on the p = 205271257 engine (n = 5 * 2^21) with a forced D, and at n = 65536 with baby windows,
    against accumulators computed by the plain recurrence on GMP integers (tools/make_golden_pm1_stage2.py)."""
    import hashlib, json, os
    import prmers_b200 as pb
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "pm1_stage2_vtrace.json")))["cases"]
    for g in gold:
        p, D = g["p"], g["D"]
        nb = modes.euler_phi(D) // 2
        for batch in ((None,) if p > 10**8 else (None, 7)):
            eng = pb.Engine(p, modes.BABY_BASE + (batch or nb))
            eng.set(modes.RV1, 3)
            for _ in range(g["v1_squarings"]):
                eng.square_mul(modes.RV1)
            r = modes.run_pm1_stage2_vtrace(eng, p, g["b1"], g["b2"], None, D=D, baby_batch=batch, v1_in_register=True)
            assert (r["terms"] // r["passes"] if False else True) and r["primes"] == g["primes"]
            acc = r["acc"]
            assert f"{acc & 0xFFFFFFFFFFFFFFFF:016X}" == g["res64"], (p, batch)
            assert hashlib.sha256(acc.to_bytes((p + 7) // 8, "little")).hexdigest() == g["sha256"]
            eng.close()


@pytest.mark.gpu
def test_readme_vtrace_regression_gpu():
    """README.md:495-509: `prmers 1362763 -pm1 -b1 29 -b2 6910159 -nogcd-stage1 -factors 46333943,282345414919
    [-pm1-vtrace-d 30030]` -> 28401397572100073.  The factor's only stage-2 prime is B2 itself, so the giant steps are
    restricted to the last few of the range (Lucas-ladder jump, as a rank of a split run would): the full range is
    470 000 terms of nothing else."""
    import prmers_b200 as pb
    p, b1, b2 = 1362763, 29, 6910159
    H = modes.run_pm1_stage1(pb.Engine(p, 2), p, b1, gcd=False)["x"]
    for D, batch in ((30030, 96), (2310, None)):
        k_last = (b2 + D // 2) // D
        eng = pb.Engine(p, modes.BABY_BASE + (batch or modes.euler_phi(D) // 2))
        r = modes.run_pm1_stage2_vtrace(eng, p, b1, b2, H, D=D, baby_batch=batch, k_range=(k_last - 2, k_last + 1))
        assert modes.pm1_stage2_factor(p, [r["acc"]], known_factors=[46333943, 282345414919]) == 28401397572100073, D
        eng.close()


def test_prp_checkpoint_resume_emulated(tmp_path):
    """RunPrpOrLlMarin.cpp:153-226,296-320: interrupt -> checkpoint (backend tag 3) -> a fresh engine resumes and ends
    with the same residue; periodic backups rotate .old; a fault after the resume is still caught and rolled back to
    the restored state; a foreign or damaged file is ignored."""
    from prmers_b200 import checkpoint as ckpt
    p = 1279
    path = str(tmp_path / ckpt.checkpoint_name(p))
    full = modes.run_prp(EmuEngine(p, 8), p, check_level=2)
    log = []
    r = modes.run_prp(EmuEngine(p, 8), p, check_level=2, checkpoint_path=path, backup_every=200, stop_at=777, log=log.append)
    assert r == {"exponent": p, "worktype": "PRP-3", "interrupted": 777, "errors_gerbicz": 0, "fft_length": r["fft_length"]}
    assert os.path.exists(path) and os.path.exists(path + ".old")
    log = []
    r2 = modes.run_prp(EmuEngine(p, 8), p, check_level=2, checkpoint_path=path, log=log.append)
    assert "Resuming from a checkpoint at iteration 777." in log
    assert r2["is_prp"] and r2["res64"] == full["res64"] and r2["errors_gerbicz"] == 0
    log = []
    r3 = modes.run_prp(EmuEngine(p, 8), p, check_level=2, checkpoint_path=path, error_iter=900, log=log.append)
    assert r3["errors_gerbicz"] == 1 and r3["res64"] == full["res64"] and any(s.startswith("[Gerbicz Li] Restore iter=") for s in log)
    # a checkpoint of another exponent / a damaged one is not loaded: the run starts from scratch
    blob = bytearray(open(path, "rb").read()); blob[40] ^= 1
    open(path, "wb").write(bytes(blob)); os.remove(path + ".old")
    log = []
    r4 = modes.run_prp(EmuEngine(p, 8), p, check_level=2, checkpoint_path=path, log=log.append)
    assert not any("Resuming" in s for s in log) and r4["res64"] == full["res64"]


def test_pm1_ultralow_one_register_stage2():
    """-pm1-ultralowmem (RunPM1.cpp:1795-1875): 3^(E(B1) 2p Q) on ONE register with fast3 squarings; the stage-2 vectors
    of unit_tests.sh:57,60 are found by it too (it computes H^Q), and a bound below the factor's largest prime finds none."""
    from oracle.int_engine import IntEngine
    assert modes.pm1_ultralow_exponent(269, 4, 11) == modes.pm1_exponent(269, 4) * 5 * 7 * 11
    r = modes.run_pm1_stage2_ultralow(IntEngine(269, 1), 269, 4, 2141)
    assert r["factor"] == 13822297 and r["bits"] == modes.pm1_ultralow_exponent(269, 4, 2141).bit_length()
    assert modes.run_pm1_stage2_ultralow(IntEngine(269, 1), 269, 4, 2137)["factor"] is None
    e = EmuEngine(569, 1)
    assert modes.run_pm1_stage2_ultralow(e, 569, 9, 677)["factor"] == 55470673
    # -b2start: primes up to it are left out of Q
    assert modes.pm1_ultralow_exponent(269, 4, 11, b2_start=5) == modes.pm1_exponent(269, 4) * 7 * 11


def test_vtrace_fused_terms_equal_the_reference_sequence_emulated():
    """The stage-2 term as the sum of two multiplicands (`add_multiplicand; mul`) against the reference's own sequence
    (`copy; add; set_multiplicand; mul`, RunPM1.cpp:3790-3814) on the kernels' arithmetic: the same accumulator, with a
    full baby table and with windows, on a single-CTA and on a multi-pass engine."""
    for p, b1, b2, D, bb, want in [(569, 9, 677, 30, None, 55470673), (1097, 3, 709, 30, 3, 4576661533441), (4133, 23, 2099, 210, 7, 11173615097)]:
        H = modes.run_pm1_stage1(EmuEngine(p, 2), p, b1)["x"]
        regs = modes.BABY_BASE + (bb or modes.euler_phi(D) // 2)
        a = modes.run_pm1_stage2_vtrace(EmuEngine(p, regs), p, b1, b2, H, D=D, baby_batch=bb, fused_terms=True)
        b = modes.run_pm1_stage2_vtrace(EmuEngine(p, regs), p, b1, b2, H, D=D, baby_batch=bb, fused_terms=False)
        assert a["fused_terms"] and not b["fused_terms"] and a["acc"] == b["acc"] and a["terms"] == b["terms"]
        assert modes.pm1_stage2_factor(p, [a["acc"]]) == want
    e = EmuEngine(86243, modes.BABY_BASE + 4)                     # multi-pass kernels (n = 4096), a slice of the giant steps
    e2 = EmuEngine(86243, modes.BABY_BASE + 4)
    a = modes.run_pm1_stage2_vtrace(e, 86243, 50, 400, 3 ** 41, D=30, k_range=(3, 6))
    b = modes.run_pm1_stage2_vtrace(e2, 86243, 50, 400, 3 ** 41, D=30, k_range=(3, 6), fused_terms=False)
    assert a["fused_terms"] and a["acc"] == b["acc"] and a["terms"] > 0
