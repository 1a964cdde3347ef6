"""ECM twisted-Edwards driver (prmers_b200/ecm.py, SURVEY 8 f4 / BASELINE config 5).

CPU tier: seed mixing, scalars, curve generation, and the register schedules run on the kernel emulator against
the big-integer engine (oracle/int_engine.py) and the closed formulas.  GPU tier: the same on the CUDA engine,
against tests/golden/ecm_*.json (tools/make_golden_ecm.py)."""
import hashlib
import json
import os

import pytest

from oracle.int_engine import IntEngine
from prmers_b200 import ecm
from tests.emu.emu import EmuEngine

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def sha(x):
    return None if x is None else hashlib.sha256(hex(x).encode()).hexdigest()[:32]


def test_splitmix64_known_answers():
    # published splitmix64 test vector (seed 0)
    s, z0 = ecm.splitmix64_step(0)
    s, z1 = ecm.splitmix64_step(s)
    assert (z0, z1) == (0xE220A8397B1DCDAF, 0x6E789E6AA1B965F4)
    assert ecm.mix64(123456789, 0) == 2600022760115733002 and ecm.mix64(123456789, 0) != ecm.mix64(123456789, 1)
    N = (1 << 701) - 1
    assert 2 < ecm.rnd_mpz_bits(N, 5, 63) < 1 << 64 and ecm.rnd_mpz_bits(N, 5, 256).bit_length() > 200


def test_scalars():
    assert ecm.stage1_exponent(10) == 2520 and ecm.stage1_exponent(1) == 1
    assert ecm.ecm_primes(10, 30) == ([2, 3, 5, 7], [11, 13, 17, 19, 23, 29])
    for e in (1, 2, 3, 7, 11, 2520, 3 ** 40 + 12345, ecm.stage1_exponent(300)):
        d = ecm.naf(e)
        assert sum(di << i for i, di in enumerate(d)) == e and d[-1] != 0
        assert all(not (d[i] and d[i + 1]) for i in range(len(d) - 1))
    assert [ecm.s2_chunk_bits(n) for n in (32, 256, 1024, 4096, 16384, 1 << 23)] == [1048576, 1048576, 524288, 262144, 131072, 65536]
    ps = ecm.ecm_primes(100, 2000)[1]
    ch = ecm.s2_chunks(ps, 64)
    assert ch[0][0] == 0 and ch[-1][1] == len(ps) and all(a[1] == b[0] for a, b in zip(ch, ch[1:]))
    assert all(e.bit_length() >= 64 for _, _, e in ch[:-1])


def test_curve_from_seed_is_on_curve():
    for p in (101, 701, 1277):
        N = (1 << p) - 1
        cv = ecm.curve_from_seed(N, ecm.mix64(123456789, 0))
        a, d, x, y = cv["a"], cv["d"], cv["X0"], cv["Y0"]
        assert (a * x * x + y * y - 1 - d * x * x * y * y) % N == 0 and a != d and d % N not in (0, 1)
    # a sigma that shares a factor with N gives the factor straight away (invm returning 1)
    with pytest.raises(ecm.FactorFound) as f:
        ecm.curve_from_sigma((1 << 101) - 1, 7432339208719)
    assert f.value.g == 7432339208719


@pytest.mark.parametrize("p,seed", [(127, 7), (101, 123456789), (521, 3)])
def test_point_schedules_match_big_integers_emulated(p, seed):
    """Every register schedule (doubling, +P, -P) on the emulated kernels reproduces the closed formulas."""
    N = (1 << p) - 1
    cv = ecm.curve_from_seed(N, seed)
    K = ecm.stage1_exponent(40)
    digits = ecm.naf(K)
    assert -1 in digits and 1 in digits
    want = ecm.reference_scalar_mul(N, cv["a"], cv["d"], cv["X0"], cv["Y0"], digits)
    for eng in (EmuEngine(p, ecm.REGS), IntEngine(p, ecm.REGS)):
        r = ecm.run_curve(eng, p, 40, 0, seed, curve=cv, K=K)
        got = tuple(ecm._value(eng, reg, p) for reg in (3, 4, 1, 5))
        if r["status"] == "NF":                       # no factor: the point registers still hold k P
            assert got == tuple(v % N for v in want)
        assert r["stage1_T"] == want[3] % N


def test_notwist_schedules_emulated():
    """The a = 1 schedules (torsion-16 curves of the reference) against the same formulas with a = 1."""
    p = 127; N = (1 << p) - 1
    x, y = 1234567, 7654321
    inv = pow(x * x * y * y, -1, N)
    d = (x * x + y * y - 1) * inv % N                 # force_on_curve (:2173-2186) with a = 1
    digits = ecm.naf(ecm.stage1_exponent(30))
    want = ecm.reference_scalar_mul(N, 1, d, x, y, digits)
    eng = EmuEngine(p, ecm.REGS)
    ecm._set_int(eng, 16, 1, p); ecm._set_int(eng, 29, d, p); eng.set(8, 2); eng.set(1, 1)
    ecm._load_base(eng, p, N, x, y)
    sgn = -1 if digits[-1] < 0 else 1
    for r, v in ((3, sgn * x % N), (4, y), (5, sgn * x * y % N)):
        ecm._set_int(eng, r, v, p)
    ecm._scalar_mul(eng, ecm.EdwardsOps(eng, notwist=True), digits)
    assert tuple(ecm._value(eng, r, p) for r in (3, 4, 1, 5)) == tuple(v % N for v in want)


def test_run_ecm_m101_emulated_matches_golden():
    gold = json.load(open(os.path.join(GOLD, "ecm_m101.json")))
    log = []
    r = ecm.run_ecm(lambda p, regs: EmuEngine(p, regs), 101, 300, 2000, curves=3, seed=123456789, log=log.append)
    assert r["status"] == "found" and str(r["factor"]) == gold["factor"] == "7432339208719"
    assert r["curves_tested_for_found"] == 1 and len(r["curves"]) == 1      # stops at the first new factor
    c, g = r["curves"][0], gold["per_curve"][0]
    assert (str(c["curve_seed"]), c["stage"], sha(c["stage1_T"]), sha(c["stage2_T"])) == (g["curve_seed"], g["stage"], g["stage1_T_sha"], g["stage2_T_sha"])
    assert "[ECM] check invariant OK" in log
    # a known factor is reported and skipped, the loop goes on (is_known, :953-962)
    r = ecm.run_ecm(lambda p, regs: EmuEngine(p, regs), 101, 300, 2000, curves=2, seed=123456789, known_factors=[7432339208719])
    assert r["status"] == "NF" and len(r["curves"]) == 2


def test_recorded_point_operations_equal_direct_calls_emulated():
    """The three point operations recorded once and replayed (k_program on the GPU, its emulation here) give the same
    residues in every register as the same schedules issued op by op."""
    p, seed = 701, 5
    K = ecm.stage1_exponent(120)
    a, b = EmuEngine(p, ecm.REGS), EmuEngine(p, ecm.REGS)
    ra = ecm.run_curve(a, p, 120, 400, seed, K=K, graphs=True)
    rb = ecm.run_curve(b, p, 120, 400, seed, K=K, graphs=False)
    assert {k: ra[k] for k in ("status", "stage", "factor", "stage1_T", "point_ops")} == {k: rb[k] for k in ("status", "stage", "factor", "stage1_T", "point_ops")}
    for r in range(ecm.REGS):
        if r not in (11, 43, 44, 45, 46, 50):           # multiplicand registers hold transform-domain data: compare raw digits only where digits exist
            assert (a.get_digits(r) == b.get_digits(r)).all(), r


def test_recording_rules_emulated():
    e = EmuEngine(701, 8)
    with pytest.raises(RuntimeError):
        e.graph_end()                                    # nothing being recorded
    e.set(0, 3); e.set(1, 5)
    e.graph_begin()
    e.square_mul(0, 3); e.set_multiplicand(2, 1); e.mul(0, 2); e.add(0, 1); e.sub_reg(0, 1); e.copy(3, 0); e.sub(3, 2); e.set(4, 9)
    with pytest.raises(RuntimeError):
        e.get_digits(0)                                  # blocking calls are refused while recording
    g = e.graph_end()
    N = (1 << 701) - 1
    val = lambda r: ecm._value(e, r, 701)
    assert val(0) == 3                                    # recording ran nothing
    x = 3
    for _ in range(10):
        e.graph_launch(g); x = x * x * 3 * 5 % N
    assert val(0) == x and val(3) == (x - 2) % N and val(4) == 9
    with pytest.raises(RuntimeError):
        EmuEngine(86243, 2).graph_begin()                 # multi-pass engines cannot be recorded


def test_invariant_catches_injected_error_emulated():
    """-erroriter (:3285-3288): Z -= 2 after some step must break a X^2 + Y^2 = Z^2 + d T^2."""
    log = []
    r = ecm.run_curve(EmuEngine(127, ecm.REGS), 127, 60, 0, 99, error_iter=17, log=log.append)
    assert r["status"] == "invariant" and "Injected error at iteration 17" in log and "[ECM] invariant FAIL" in log


def test_curve_sharding_covers_all_curves():
    from prmers_b200 import multi
    for world in (1, 2, 3, 8):
        got = sorted(c for r in range(world) for c in multi.assign(list(range(8)), r, world))
        assert got == list(range(8))


def _check_against_golden(name, engine_factory, graphs):
    gold = json.load(open(os.path.join(GOLD, name + ".json")))
    r = ecm.run_ecm(engine_factory, gold["p"], gold["b1"], gold["b2"], curves=gold["curves"], seed=gold["seed"],
                    continue_after_factor=True, graphs=graphs)
    assert r["status"] == gold["status"] and (None if r["factor"] is None else str(r["factor"])) == gold["factor"]
    assert len(r["curves"]) == len(gold["per_curve"])
    for c, g in zip(r["curves"], gold["per_curve"]):
        assert str(c["curve_seed"]) == g["curve_seed"] and c["status"] == g["status"] and c["stage"] == g["stage"]
        assert (None if c["factor"] is None else str(c["factor"])) == g["factor"]
        assert sha(c.get("stage1_T")) == g["stage1_T_sha"] and sha(c.get("stage2_T")) == g["stage2_T_sha"]
        assert c.get("point_ops") == g["point_ops"]
    return r


@pytest.mark.gpu
@pytest.mark.parametrize("graphs", [True, False])
def test_ecm_m701_config5_gpu(graphs):
    """BASELINE config 5: M701, B1 = 6000, B2 = 33333, 8 curves, fixed seed -- every curve's residue and factor."""
    import prmers_b200 as pb
    r = _check_against_golden("ecm_m701", lambda p, regs: pb.Engine(p, regs), graphs)
    assert r["fft_length"] == 32


@pytest.mark.gpu
def test_ecm_stage2_gpu():
    """Curves that run stage 2 to the end (M1277 has no small factor; M101 finds its factor in stage 2)."""
    import prmers_b200 as pb
    _check_against_golden("ecm_m1277", lambda p, regs: pb.Engine(p, regs), True)
    _check_against_golden("ecm_m101", lambda p, regs: pb.Engine(p, regs), True)


@pytest.mark.gpu
@pytest.mark.parametrize("p,force_s", [(86243, 64), (216091, 64)])
def test_ecm_point_ops_on_the_multi_pass_path_gpu(p, force_s):
    """addsub / mul_copy / sub_reg on 3-pass engines (n = 4096 forced off the single-CTA path, and n = 10240 = 5 * 2^11):
    the Edwards schedules against big integers."""
    import prmers_b200 as pb
    N = (1 << p) - 1
    cv = ecm.curve_from_seed(N, 2024)
    K = ecm.stage1_exponent(60)
    want = ecm.reference_scalar_mul(N, cv["a"], cv["d"], cv["X0"], cv["Y0"], ecm.naf(K))
    eng = pb.Engine(p, ecm.REGS, force_S=force_s)
    assert not eng.is_single_cta()
    r = ecm.run_curve(eng, p, 60, 0, 2024, curve=cv, K=K)
    assert r["status"] == "NF" and r["stage1_T"] == want[3] % N
    assert tuple(eng.get_mpz(reg) for reg in (3, 4, 1)) == tuple(v % N for v in want[:3])


@pytest.mark.gpu
def test_graph_replay_equals_direct_calls_gpu():
    import numpy as np
    import prmers_b200 as pb
    e = pb.Engine(701, 8)
    with pytest.raises(RuntimeError):
        e.graph_end()
    e.set(0, 3); e.set(1, 5)
    e.graph_begin()
    e.square_mul(0, 3); e.set_multiplicand(2, 1); e.mul(0, 2); e.add(0, 1); e.sub_reg(0, 1); e.copy(3, 0)
    with pytest.raises(RuntimeError):
        e.get(0)                                     # blocking calls are refused while recording
    g = e.graph_end()
    x, N = 3, (1 << 701) - 1
    assert e.get_mpz(0) == 3                          # recording did not run anything
    for _ in range(50):
        e.graph_launch(g)
        x = x * x * 3 * 5 % N
    assert e.get_mpz(0) == x and e.get_mpz(3) == x
    big = pb.Engine(1362763, 2)
    with pytest.raises(RuntimeError):
        big.graph_begin()                             # multi-pass engines carry per-launch epochs
    assert np.array_equal(e.get(0), e.get(3))


def test_ecm_stage1_resume_line(tmp_path):
    """A stage 1 without a factor appends the GMP-ECM resume line (RunEcmTwistedEdwards.cpp:1287-1330, 3396-3426): the
    Montgomery x of K P and the coefficient A of the curve, against the big-integer double-and-add on the Edwards curve
    (u = (1 + y) / (1 - y), A = 2 (a + d) / (a - d)); the line's checksum is the reference's."""
    p, b1 = 701, 50
    N = (1 << p) - 1
    seed = ecm.mix64(7, 1)                                       # this curve ends stage 1 without a factor
    path = str(tmp_path / ecm.ecm_resume_name(p, b1))
    r = ecm.run_curve(EmuEngine(p, ecm.REGS), p, b1, 0, seed, resume_path=path)
    assert r["status"] == "NF" and "resume_x" in r
    cv = ecm.curve_from_seed(N, seed)
    X, Y, Z, T = ecm.reference_scalar_mul(N, cv["a"], cv["d"], cv["X0"], cv["Y0"], ecm.naf(ecm.stage1_exponent(b1)))
    assert r["resume_x"] == (Z + Y) * pow((Z - Y) % N, -1, N) % N
    assert r["resume_A"] == 2 * (cv["a"] + cv["d"]) * pow((cv["a"] - cv["d"]) % N, -1, N) % N
    got = ecm.parse_ecm_resume_line(open(path).read().splitlines()[-1])
    assert got == {"p": p, "b1": b1, "x": r["resume_x"], "A": r["resume_A"]}
    # the point lies on the Montgomery curve B v^2 = u^3 + A u^2 + u for SOME B: u (u^2 + A u + 1) is B times a square;
    # cheaper and sharper: the same u comes out of the affine Edwards point
    y = Y * pow(Z, -1, N) % N
    assert r["resume_x"] == (1 + y) * pow((1 - y) % N, -1, N) % N
    with pytest.raises(ValueError):
        ecm.parse_ecm_resume_line(open(path).read().replace("B1=50", "B1=51"))
    # the curve loop appends one line per curve that ends stage 1 without a factor
    path2 = str(tmp_path / "loop.save")
    out = ecm.run_ecm(lambda q, regs: EmuEngine(q, regs), p, b1, 0, curves=3, seed=7, continue_after_factor=True, resume_path=path2)
    lines = open(path2).read().splitlines()
    assert len(lines) == sum(1 for c in out["curves"] if "resume_x" in c) >= 1
    assert [ecm.parse_ecm_resume_line(l)["x"] for l in lines] == [c["resume_x"] for c in out["curves"] if "resume_x" in c]


def _random_program(e, rng, regs, count):
    """A hazard-rich op list on registers 0 .. regs-1 (register regs-1 is kept as a multiplicand slot)."""
    M = regs - 1
    for _ in range(count):
        a, b = rng.randrange(M), rng.randrange(M)
        k = rng.randrange(9)
        if k == 0:
            e.copy(a, b)
        elif k == 1:
            e.square_mul(a, rng.choice([1, 3]))
        elif k == 2:
            e.set_multiplicand(M, b); e.mul(a, M)
        elif k == 3:
            e.add(a, b)
        elif k == 4:
            e.sub_reg(a, b)
        elif k == 5:
            e.sub(a, rng.randrange(1, 50))
        elif k == 6:
            e.set(a, rng.randrange(2, 1000))
        elif k == 7:
            s_, d_, x_, y_ = a, (a + 1) % M, b, (b + 2) % M      # outputs apart from inputs: a recording stores the reference's
            e.addsub(s_, d_, x_, y_) if not ({s_, d_} & {x_, y_}) else e.add(a, b)   # default composition (engine.h:72-79)
        else:
            e.mul_copy(a, M, b) if a != b else e.mul(a, M)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_level_schedule_of_recorded_programs_emulated(seed):
    """Recorded op sequences are stored as list schedules over the warp groups of the CTA (schedule_program), ordered by
    per-op wait counts only; the emulator plays the groups in an order as far from the recorded one as the counts allow.
    Registers after the replay = registers after the same calls made directly, for random op lists full of
    read-after-write, write-after-read and write-after-write pairs."""
    import random
    p, regs = 701, 7
    m = (1 << p) - 1
    a, b = EmuEngine(p, regs), EmuEngine(p, regs)
    rng = random.Random(seed)
    for r in range(regs):
        v = rng.randrange(m)
        a.set_mpz(r, v); b.set_mpz(r, v)
    a.set_multiplicand(regs - 1, 0); b.set_multiplicand(regs - 1, 0)
    b.graph_begin()
    _random_program(b, random.Random(100 + seed), regs, 60)
    g = b.graph_end()
    for _ in range(2):
        _random_program(a, random.Random(100 + seed), regs, 60)
        b.graph_launch(g)
    for r in range(regs - 1):
        assert a.get_mpz(r) == b.get_mpz(r), r


def test_level_schedule_shapes_of_the_edwards_programs():
    """What the schedule buys: the doubling (23 recorded ops) and the mixed additions (40) of the twisted-Edwards driver as
    list schedules over four warp groups, with the scratch multiplicand register renamed per lifetime.  Length in the
    scheduler's cost model (0.1 us; a transform with carry 45, a forward transform 25, a linear op 10): 590 and 975 when
    the ops run one after the other."""
    e = EmuEngine(701, ecm.REGS)
    ops = ecm.EdwardsOps(e, graphs=True)
    shapes = {k: e.graph_shape(g) for k, g in ops._g.items()}
    assert shapes["dbl"][1] == 23 and shapes["add+"][1] == 40 and shapes["add-"][1] == 40, shapes
    assert shapes["dbl"][0] <= 200 and shapes["add+"][0] <= 210 and shapes["add-"][0] <= 210, shapes


def test_long_programs_emulated():
    """Programs longer than the shared-memory staging area of k_program (128 ops) and longer than the scheduler's limit
    (4096 ops: recorded order, one group) replay like the direct calls."""
    import random
    p, regs = 701, 7
    m = (1 << p) - 1
    for count in (300, 3000):
        a, b = EmuEngine(p, regs), EmuEngine(p, regs)
        rng = random.Random(count)
        for r in range(regs):
            v = rng.randrange(m)
            a.set_mpz(r, v); b.set_mpz(r, v)
        a.set_multiplicand(regs - 1, 0); b.set_multiplicand(regs - 1, 0)
        b.graph_begin()
        _random_program(b, random.Random(5), regs, count)
        g = b.graph_end()
        assert b.graph_shape(g)[1] >= count
        _random_program(a, random.Random(5), regs, count)
        b.graph_launch(g)
        for r in range(regs - 1):
            assert a.get_mpz(r) == b.get_mpz(r), (count, r)


@pytest.mark.gpu
@pytest.mark.parametrize("p,batches,count", [(701, (1, 3), 80), (9941, (1, 3), 80), (701, (1, 2500), 300), (40009, (1,), 3000)])
def test_level_schedule_of_recorded_programs_gpu(p, batches, count):
    """The same on the device (n = 32: four warp groups of one warp; n = 512: four of 128 threads; n = 2048: one of 512),
    batched and not; 2500 slabs do not fit the GPU in one wave with four groups per CTA, so the launch shares the
    schedule's groups among fewer (k_program: group of the schedule modulo groups of the launch); 300 ops are read from
    global memory instead of the staging area, 3000 recorded calls (> 4096 ops) run in recorded order on one group."""
    import random
    import prmers_b200 as pb
    regs = 7
    m = (1 << p) - 1
    for batch in batches:
        a, b = pb.Engine(p, regs, batch=batch), pb.Engine(p, regs, batch=batch)
        rng = random.Random(7)
        for s in range(batch):
            for r in range(regs):
                v = rng.randrange(m)
                (a.slab(s) if batch > 1 else a).set_mpz(r, v); (b.slab(s) if batch > 1 else b).set_mpz(r, v)
        a.set_multiplicand(regs - 1, 0); b.set_multiplicand(regs - 1, 0)
        b.graph_begin()
        _random_program(b, random.Random(11), regs, count)
        g = b.graph_end()
        for _ in range(3 if count <= 300 else 1):
            _random_program(a, random.Random(11), regs, count)
            b.graph_launch(g)
        for s in (range(batch) if batch <= 8 else (0, 1, batch // 2, batch - 1)):
            va, vb = (a.slab(s), b.slab(s)) if batch > 1 else (a, b)
            for r in range(regs - 1):
                assert va.get_mpz(r) == vb.get_mpz(r), (batch, s, r)
        a.close(); b.close()


# ---- batched engines: all curves of a call in lockstep, one CTA per curve (round 2) --------------------------------------
def _curve_key(c):
    return (c["curve"], c["curve_seed"], c["status"], c["factor"], c["stage"], c.get("stage1_T"), c.get("stage2_T"))


def test_batched_engine_ops_emulated():
    """Every register op of a batched engine acts on each slab exactly as a plain engine would (emulated kernels)."""
    import random
    p, B = 701, 3
    m = (1 << p) - 1
    e = EmuEngine(p, 12, batch=B)
    rng = random.Random(1)
    xs, ys = [rng.randrange(m) for _ in range(B)], [rng.randrange(m) for _ in range(B)]
    for s in range(B):
        e.set_mpz(0, xs[s], slab=s); e.set_mpz(1, ys[s], slab=s)
    e.set_multiplicand(2, 1); e.mul(0, 2, 3); e.square_mul(0); e.add(0, 1); e.sub_reg(0, 1); e.sub(0, 2); e.copy(3, 0)
    e.addsub(4, 5, 0, 1); e.set(6, 7); e.mul_add(3, 2, 6)
    e.graph_begin(); e.square_mul(0); e.mul(0, 2); e.add(0, 6); g = e.graph_end()
    want = []
    for s in range(B):
        x = xs[s] * ys[s] * 3 % m; x = (x * x - 2) % m
        assert e.get_mpz(0, slab=s) == x and e.get_mpz(4, slab=s) == (x + ys[s]) % m and e.get_mpz(5, slab=s) == (x - ys[s]) % m
        assert e.get_mpz(3, slab=s) == (x * ys[s] + 7) % m
        want.append(x)
    e.graph_launch(g); e.graph_launch(g)
    for s in range(B):
        x = want[s]
        for _ in range(2):
            x = (x * x * ys[s] + 7) % m
        assert e.get_mpz(0, slab=s) == x
    with pytest.raises(ValueError):
        EmuEngine(86243, 4, batch=2)                         # multi-pass sizes cannot be batched


def test_batched_ecm_equals_sequential_emulated():
    """run_ecm_batched: every curve's seed, status, stage, factor and T residues equal the sequential run's."""
    p, b1, b2, K, seed = 701, 300, 900, 5, 123456789
    seq = ecm.run_ecm(lambda q, r: EmuEngine(q, r), p, b1, b2, curves=K, seed=seed, continue_after_factor=True)
    bat = ecm.run_ecm_batched(lambda q, r, b: EmuEngine(q, r, batch=b), p, b1, b2, curves=K, seed=seed)
    assert [_curve_key(c) for c in seq["curves"]] == [_curve_key(c) for c in bat["curves"]]
    assert (seq["status"], seq["factor"]) == (bat["status"], bat["factor"]) and bat["batch"] == K
    one = ecm.run_ecm_batched(lambda q, r, b: EmuEngine(q, r, batch=b), p, b1, b2, curves=K, seed=seed, curve_indices=[3])
    assert _curve_key(one["curves"][0]) == _curve_key(seq["curves"][3])


@pytest.mark.gpu
def test_batched_ecm_config5_gpu():
    """BASELINE config 5 (M701, B1 = 6000, B2 = 33333, K = 8, seed 123456789) as ONE batched run against the golden
    per-curve records, and 64 curves at once against the sequential driver on a sample of them."""
    import prmers_b200 as pb
    gold = json.load(open(os.path.join(GOLD, "ecm_m701.json")))
    p, b1, b2, K, seed = gold["p"], gold["b1"], gold["b2"], gold["curves"], gold["seed"]
    bat = ecm.run_ecm_batched(lambda q, r, b: pb.Engine(q, r, batch=b), p, b1, b2, curves=K, seed=seed)
    assert len(bat["curves"]) == len(gold["per_curve"]) == 8
    for c, g in zip(bat["curves"], gold["per_curve"]):
        assert (str(c["curve_seed"]), c["status"], c["stage"]) == (g["curve_seed"], g["status"], g["stage"])
        assert (None if c["factor"] is None else str(c["factor"])) == g["factor"]
        assert sha(c.get("stage1_T")) == g["stage1_T_sha"] and sha(c.get("stage2_T")) == g["stage2_T_sha"]
    big = ecm.run_ecm_batched(lambda q, r, b: pb.Engine(q, r, batch=b), p, 600, 2000, curves=64, seed=7)
    for ci in (0, 17, 63):
        seq = ecm.run_ecm(lambda q, r: pb.Engine(q, r), p, 600, 2000, curves=64, seed=7, continue_after_factor=True, curve_indices=[ci])
        assert _curve_key(seq["curves"][0]) == _curve_key(big["curves"][ci])


# ---- the other curve families: -torsion16 and -iv163 (round 2) ----------------------------------------------------------------
def _edwards_group_order(q, a, d):
    """Order of the group of a x^2 + y^2 = 1 + d x^2 y^2 over F_q (q a small prime), counted on the birationally
    equivalent Montgomery curve B v^2 = u^3 + A u^2 + u, A = 2 (a + d) / (a - d), B = 4 / (a - d)."""
    inv = pow((a - d) % q, -1, q)
    A, Binv = 2 * (a + d) * inv % q, (a - d) * pow(4, -1, q) % q
    sq = {x * x % q for x in range(1, q)}
    n = 1
    for u in range(q):
        rhs = (u * u * u + A * u * u + u) * Binv % q
        n += 1 if rhs == 0 else (2 if rhs in sq else 0)
    return n


def test_torsion16_and_family_iv_curves():
    """Generators of RunEcmTwistedEdwards.cpp:2228-2435: the curves they return satisfy the curve equation with the
    advertised a (1, resp. -1), and have the advertised torsion: reduced modulo small primes their group order is always
    divisible by 16 (-torsion16: Z/2 x Z/8) resp. by 8 at least (family IV)."""
    for p in (701, 1277):
        N = (1 << p) - 1
        for c in range(4):
            cs = ecm.mix64(987654321, c)
            for tor, a_want in (("16", 1), ("family_iv_163", N - 1)):
                cv = ecm.curve_for(N, cs, tor)
                assert cv is not None and cv["a"] == a_want and cv["torsion"] == tor
                x2, y2 = cv["X0"] ** 2 % N, cv["Y0"] ** 2 % N
                assert (cv["a"] * x2 + y2 - 1 - cv["d"] * x2 * y2) % N == 0
    checked = 0
    for q in (10007, 10009, 20011, 30011):
        for c in range(6):
            cs = ecm.mix64(55, c)
            for tor, div in (("16", 16), ("family_iv_163", 8)):
                try:
                    cv = ecm.curve_for(q, cs, tor)
                except ecm.FactorFound:
                    continue
                if cv is None:
                    continue
                a, d = cv["a"] % q, cv["d"] % q
                if d in (0, a):
                    continue
                assert _edwards_group_order(q, a, d) % div == 0, (q, c, tor)
                checked += 1
    assert checked >= 6
    with pytest.raises(ValueError):
        ecm.curve_for((1 << 101) - 1, 1, "bogus")


def test_ecm_finds_factors_with_every_curve_family():
    """The a = 1 ("notwist") schedules with -torsion16 curves and the general-a schedules with family-IV curves on exact
    arithmetic: invariant holds on every curve, and M67 = 193707721 * 761838257287 is split."""
    for tor in ("16", "family_iv_163", None):
        r = ecm.run_ecm(lambda p, regs: IntEngine(p, regs), 67, 4000, 40000, curves=12, seed=11, continue_after_factor=True, torsion=tor)
        assert all(c["status"] != "invariant" for c in r["curves"])
        assert r["status"] == "found" and ((1 << 67) - 1) % r["factor"] == 0 and 1 < r["factor"] < (1 << 67) - 1, tor
