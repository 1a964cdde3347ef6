#!/usr/bin/env python
"""bench.py -- PRP squaring throughput of the B200 Marin engine (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            (torchrun launches it for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is ITERS consecutive square_mul on one resident register: the PRP hot loop of
RunPrpOrLlMarin.cpp:293-321.  Rank 0 works on M136279841 (n = 2^23, the headline config); further
ranks each take their own prime exponent near 1.4e8 (same n), one engine per GPU, no data-path
collective (a squaring chain is strictly sequential: replicas only).  Timing: CUDA events on the
engine's own stream around each step, barrier + synchronise on both sides, max over ranks.

One JSON line on rank 0:
  value      whole-job iterations/s, operands resident in HBM
  e2e        same metric through the C ABI with HOST buffers: every step uploads the residue digits
             from pinned host memory (set_digits), runs ITERS squarings and reads the canonical digits back
  roofline   dominant kernel: algorithmic bytes (8n read + 8n write per pass) / CUDA-event duration
  cpu_baseline  the oracle's restatement of the Marin algorithm timed on this box's host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

HEADLINE_P = 136279841
OTHER_P = [139123459, 139246931, 139370393, 139493863, 139617329, 139740787, 139864247]  # primes, all n = 2^23
ITERS = 1000            # squarings per step
METRIC = "PRP iterations/s, M136279841-class exponents (n = 2^23 IBDWT), whole job"
# steady state of a squaring chain: two launches per square_mul (the inverse pass also runs the next
# squaring's forward strided pass on the tile it holds); (name, bench_pass id, algorithmic bytes / n)
PASSES = [("k_mid_fast<SQR> (contiguous pass: forward, pointwise square, inverse)", 1, 16),
          ("k_inv_fast_T4<true> (strided inverse + unweight/carry/reweight + next strided forward)", 3, 32)]
SUMMARY_KEYS = ["k_mid_fast", "k_inv_fast_T4"]          # kernel keys of profiles/ncu_summary.json, same order


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline(sample_iters, p=HEADLINE_P):
    """The oracle port (exact restatement of the Marin IBDWT, OpenMP over the host cores)."""
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count())      # torchrun pins it to 1; the baseline uses every core
    from oracle import marin_oracle as mo
    o = mo.OracleEngine(p, 1)
    o.set(0, 3)
    for _ in range(6):
        o.square_mul(0)                      # digits become dense after a few squarings
    t0 = time.perf_counter()
    for _ in range(sample_iters):
        o.square_mul(0)
    dt = time.perf_counter() - t0
    return {"value": sample_iters / dt, "unit": "iter/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{sample_iters} square_mul at p = {p} (n = {o.n}) after 6 warm-up squarings, "
                      f"oracle/marin_oracle.c (cache-blocked two-pass NTT, OpenMP over all host threads); the reference's OpenCL path "
                      f"needs PoCL + OpenCL / GMP headers: the GPU box has none of them (profiles/r2/probe_box.log)"}, o


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path = the oracle port (the stock binary cannot be
    built on the GPU box: no CL/cl.h, no gmp.h, no PoCL -- probed, profiles/r2/probe_box.log)."""
    if rank != 0:
        return
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count())      # torchrun pins it to 1; the reference arm uses every core
    from oracle import marin_oracle as mo
    sample = 8
    o = mo.OracleEngine(HEADLINE_P, 1)
    o.set(0, 3)
    for _ in range(4):
        o.square_mul(0)
    for _ in range(args.warmup):
        for _ in range(sample):
            o.square_mul(0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for _ in range(sample):
            o.square_mul(0)
    dt = time.perf_counter() - t0
    v = args.steps * sample / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "iter/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 (mod 2^64-2^32+1)",
            "data": "synthetic (genuine PRP sequence 3^(2^i) mod M136279841)",
            "config": {"workload": f"PRP M{HEADLINE_P}, n = 2^23 Marin IBDWT, {sample} square_mul per step (bounded sample)",
                       "exponent": HEADLINE_P, "transform_size": o.n},
            "cpu_baseline": {"value": v, "unit": "iter/s", "cores": os.cpu_count(), "kind": "port",
                             "sample": f"{sample} square_mul per step on all host threads (OpenMP), oracle/marin_oracle.c with the "
                                       f"cache-blocked two-pass NTT; no PoCL / OpenCL / GMP headers on the box (profiles/r2/probe_box.log)"},
            "e2e": {"value": v, "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_workload(args, rank, local_rank, world):
    """--workload pm1 | ecm: BASELINE configs 4 and 5, one JSON line each (the default headline line is unchanged)."""
    import torch
    import torch.distributed as dist
    import prmers_b200 as pb
    from prmers_b200 import ecm, modes, multi
    dev = torch.device("cuda", local_rank)
    factory = lambda q, regs: pb.Engine(q, regs, device=local_rank)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if args.workload == "pm1":
        # config 4: p = 205271257 (n = 5 * 2^21); stage 1 is a squaring chain (ms per square_mul); stage 2 = V-trace with a
        # forced D = 210 and all 24 baby traces resident (-pm1-vtrace-d 210 -pm1-vtrace-baby-batch 24), a bounded slice of
        # the giant steps right after B1 = 1 300 000 per rank (the ranks split the range as a real run would)
        p, b1, D, steps_k = 205271257, 1300000, 210, 6 * max(args.steps, 1)
        eng = factory(p, modes.BABY_BASE + 24)
        n = eng.get_size()
        eng.set(modes.RV1, 3)
        eng.square_mul_repeat(modes.RV1, 40)
        for _ in range(max(args.warmup, 3)):
            eng.bench_square_mul(modes.RACC, 50)
        ms_sq = eng.bench_square_mul(modes.RACC, 300) / 300
        eng.set(1, 5); eng.square_mul_repeat(1, 20); eng.set_multiplicand(2, 1); eng.sync()
        t0 = time.perf_counter()
        for _ in range(200):
            eng.mul(3, 2)
        eng.sync()
        ms_mul = (time.perf_counter() - t0) / 200 * 1e3
        k_first = (b1 + 1 + D // 2) // D
        k_lo = k_first + rank * steps_k
        b2 = (k_first + world * steps_k) * D
        launches0 = eng.launch_count()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        r = modes.run_pm1_stage2_vtrace(eng, p, b1, b2, None, D=D, k_range=(k_lo, k_lo + steps_k), v1_in_register=True)
        dt = multi.max_over_ranks(time.perf_counter() - t0, device=dev)
        launches = eng.launch_count() - launches0
        # the slice above is mostly set-up (baby table, Lucas jump, read-back of the accumulator); the cost of a term in a long
        # run is the difference between a slice of twice the length and this one
        t1 = time.perf_counter()
        r2 = modes.run_pm1_stage2_vtrace(eng, p, b1, b2 + steps_k * D, None, D=D, k_range=(k_lo, k_lo + 2 * steps_k), v1_in_register=True)
        dt2 = time.perf_counter() - t1
        steady_ms = (dt2 - dt) * 1e3 / max(r2["terms"] - r["terms"], 1) if world == 1 else None
        steady_ref_ms = None
        if world == 1 and r.get("fused_terms"):
            # the same two slices with the reference's own term sequence (copy + add + set_multiplicand + mul)
            ta = time.perf_counter()
            modes.run_pm1_stage2_vtrace(eng, p, b1, b2, None, D=D, k_range=(k_lo, k_lo + steps_k), v1_in_register=True, fused_terms=False)
            tb = time.perf_counter()
            modes.run_pm1_stage2_vtrace(eng, p, b1, b2 + steps_k * D, None, D=D, k_range=(k_lo, k_lo + 2 * steps_k), v1_in_register=True, fused_terms=False)
            steady_ref_ms = ((time.perf_counter() - tb) - (tb - ta)) * 1e3 / max(r2["terms"] - r["terms"], 1)
        terms = sum(multi.gather([r["terms"]]))
        if rank == 0:
            peak, peak_src = measured_peak()
            line = {"metric": "P-1 stage-2 terms/s, V-trace, p = 205271257 (n = 5 * 2^21), whole job", "value": terms / dt, "unit": "terms/s",
                    "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dt * 1e3 / max(args.steps, 1),
                    "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 (mod 2^64-2^32+1)",
                    "data": "synthetic (V_1 = 3^(2^40) mod Mp; bounds B1 = 1300000, a slice of the giant steps above it)",
                    "config": {"workload": "BASELINE configs[3]: P-1 p = 205271257, B1 = 1300000, V-trace stage 2 with forced D = 210, 24 resident "
                                           f"baby traces, {steps_k} giant steps per rank (baby table and Lucas jump included in the time)",
                               "exponent": p, "transform_size": n, "D": D, "baby_batch": 24, "giant_steps_per_rank": steps_k,
                               "parallelism": f"giant-step range split over {world} rank(s), no data-path collective"},
                    "terms": terms, "primes": sum(multi.gather([r["primes"]])) if world == 1 else None,
                    "ms_per_square_mul": ms_sq, "ms_per_mul": ms_mul, "ms_per_term": dt * 1e3 / max(r["terms"], 1),
                    "ms_per_term_steady": steady_ms, "steady_terms_per_s": (1e3 / steady_ms) if steady_ms else None,
                    "ms_per_term_steady_reference_sequence": steady_ref_ms,
                    "term_sequence": "add_multiplicand + mul (sum of two multiplicands)" if r.get("fused_terms") else "copy + add + set_multiplicand + mul",
                    "e2e": {"value": terms / dt, "unit": "terms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": n * 8,
                            "note": "the driver runs through the public API; the accumulator residue is read back at the end"},
                    "gpu_launches": int(launches), "clocks": sampler.stop(),
                    "roofline": {"bound": "hbm", "kernel": "mul (3 passes + multiplicand read: 56 n bytes)", "achieved": 56 * n / (ms_mul * 1e-3) / 1e9,
                                 "peak": peak, "unit": "GB/s", "frac": 56 * n / (ms_mul * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src},
                    "cpu_baseline": None}
            print(json.dumps(line), flush=True)
        eng.close()
        return
    # config 5: ECM on M701, B1 = 6000, B2 = 33333, K = 8 curves (twisted Edwards, default family), fixed seed; curve c on rank c mod world
    # K = 8 is BASELINE's count (one launch now carries all 8: one CTA per curve); K = 740 fills the GPU (5 CTAs per SM)
    p, b1, b2, K, seed = 701, 6000, 33333, 8, 123456789
    bfactory = lambda q, regs, b: pb.Engine(q, regs, device=local_rank, batch=b)
    ecm.run_ecm(factory, p, b1, b2, curves=1, seed=seed, continue_after_factor=True)      # warm-up curve (tables, graphs)
    t0 = time.perf_counter()
    seq = ecm.run_ecm(factory, p, b1, b2, curves=K, seed=seed, continue_after_factor=True, curve_indices=multi.assign(list(range(K)), rank, world))
    dt_seq = multi.max_over_ranks(time.perf_counter() - t0, device=dev)
    Kbig = 740 * world          # five CTAs of 128 threads and 94 registers per SM: the curves one wave of the GPU holds
    t0 = time.perf_counter()
    big = ecm.run_ecm_sharded(factory, p, b1, b2, curves=Kbig, seed=seed, batch_engine_factory=bfactory)
    dt_big = multi.max_over_ranks(time.perf_counter() - t0, device=dev)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    reps = max(args.steps, 1)
    res = None
    for _ in range(reps):
        res = ecm.run_ecm_sharded(factory, p, b1, b2, curves=K, seed=seed, batch_engine_factory=bfactory)
    dt = multi.max_over_ranks(time.perf_counter() - t0, device=dev)
    if rank == 0:
        pops = sum(c.get("point_ops", 0) for c in res["curves"])
        line = {"metric": "ECM curves/s, M701 B1 = 6000 B2 = 33333 (twisted Edwards), whole job", "value": K * reps / dt, "unit": "curves/s",
                "n_gpus": world, "steps": reps, "warmup": 1, "ms_per_step": dt * 1e3 / reps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "u64 (mod 2^64-2^32+1)", "data": "synthetic (fixed -seed 123456789)",
                "config": {"workload": "BASELINE configs[4]: ECM M701 B1 = 6000 B2 = 33333 K = 8 curves, n = 32, 51 registers per curve, all curves of a rank in lockstep on one batched engine (one CTA per curve)",
                           "exponent": p, "transform_size": 32, "curves": K, "parallelism": f"curve c on rank c mod {world}"},
                "point_ops_per_curve": pops / max(len(res["curves"]), 1),
                "us_per_point_op_launch": dt * 1e6 / reps / max(pops / max(len(res["curves"]), 1), 1),
                "sequential_one_engine_per_curve": {"curves_per_s": K / dt_seq, "note": "the reference's structure (RunEcmTwistedEdwards.cpp:1736-1830), point ops replayed as one launch each"},
                "saturating_batch": {"curves": Kbig, "curves_per_s": Kbig / dt_big, "factors_found": sum(1 for c in big["curves"] if c["status"] == "found")},
                "factors": sorted({str(c["factor"]) for c in res["curves"] if c.get("factor")}),
                "e2e": {"value": K * reps / dt, "unit": "curves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                        "note": "host loop through the public API, gcds on the host"},
                "gpu_launches": res.get("gpu_launches"), "clocks": sampler.stop(), "roofline": None, "cpu_baseline": None}
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--iters", type=int, default=ITERS, help="square_mul per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="prp", choices=["prp", "pm1", "ecm"],
                    help="prp = the headline (default); pm1 = BASELINE config 4 (V-trace stage 2 at n = 5 * 2^21); ecm = config 5 (M701)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import prmers_b200 as pb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if args.workload != "prp":
        run_workload(args, rank, local_rank, world)
        if world > 1:
            dist.destroy_process_group()
        return

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    p = HEADLINE_P if rank == 0 else OTHER_P[(rank - 1) % len(OTHER_P)]
    eng = pb.Engine(p, 8, device=local_rank)         # 8 registers as RunPrpOrLlMarin.cpp:213 uses
    n = eng.get_size()
    eng.set(0, 3)
    eng.square_mul_repeat(0, 96)                      # the reference's -bench warm-up (App.cpp:715-716)
    iters = args.iters
    for _ in range(max(args.warmup, 3)):
        eng.bench_square_mul(0, iters)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count()
    barrier()
    ms_total = 0.0
    for _ in range(args.steps):
        ms_total += eng.bench_square_mul(0, iters)   # CUDA events on the engine's stream, final sync
    barrier()
    launches = eng.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    total_iters = world * args.steps * iters
    value = total_iters / (ms_max * 1e-3)

    # ---- end to end through the C ABI with host buffers (pinned), H2D + D2H inside the timed region
    L = pb.load_library()
    import ctypes as C
    hin = L.mb200_host_alloc(n * 8)
    hout = L.mb200_host_alloc(n * 8)
    ain = np.ctypeslib.as_array(C.cast(hin, C.POINTER(C.c_uint64)), shape=(n,))
    aout = np.ctypeslib.as_array(C.cast(hout, C.POINTER(C.c_uint64)), shape=(n,))
    eng.get(0, out=ain)                               # a genuine canonical residue as the step input
    for _ in range(2):
        eng.set_digits(0, ain); eng.square_mul_repeat(0, iters); eng.get(0, out=aout)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.set_digits(0, ain)
        eng.square_mul_repeat(0, iters)
        eng.get(0, out=aout)
        ain, aout = aout, ain                         # the next step uploads what this one downloaded (both buffers are pinned)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = total_iters / float(t.item())

    # ---- the PRP loop itself with Gerbicz-Li on (BASELINE configs[1] says "with Gerbicz checks"): modes.run_prp through
    # the API until the first check has passed, a check every 2 blocks of B = floor(sqrt p) squarings -- 25x more often
    # than the reference's default (one per 51 blocks, RunPrpOrLlMarin.cpp:257-268), so this is the pessimistic figure
    gerbicz = None
    if rank == 0:
        from prmers_b200 import modes
        g_eng = pb.Engine(p, 8, device=local_rank)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = modes.run_prp(g_eng, p, check_level=2, stop_after_checks=1)
        g_eng.sync()
        dt = time.perf_counter() - t0
        B = int(p ** 0.5)
        gerbicz = {"value": r["stopped_at"] / dt, "unit": "iter/s", "iterations": r["stopped_at"], "seconds": dt,
                   "block": B, "check_every_blocks": 2, "checks_passed": r["checks_passed"], "errors": r["errors_gerbicz"],
                   "projected_at_reference_check_interval": value / world / (1.0 + 1.0 / 51 + 3.0 / B),
                   "note": "modes.run_prp (square_mul; per block copy + set_multiplicand + mul; check = B squarings + device is_equal)"}
        g_eng.close()

    if rank == 0:
        peak, peak_src = measured_peak()
        # per-kernel durations, live, with CUDA events on the launching stream
        passes = []
        # ... on register 0: the live, dense residue (the data-dependent carry absorption runs as in the real loop)
        for _, which, _ in PASSES:
            cnt = 100
            passes.append(eng.bench_pass(0, which, cnt) / cnt)            # ms per launch
        dom = max(range(len(PASSES)), key=lambda i: passes[i])
        alg_bytes = PASSES[dom][2] * n                                    # 8n read + 8n write per pass of the 3-pass count
        achieved = alg_bytes / (passes[dom] * 1e-3) / 1e9
        traffic, alu_instr = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_summary.json")) as f:
                summ = json.load(f)
            traffic = summ["dram_bytes_per_launch"].get(SUMMARY_KEYS[dom])
            alu_instr = sum(summ["alu_pipe_warp_instructions_per_launch"].values())
        except Exception:
            pass
        ms_iter = ms_max / (args.steps * iters)
        line = {
            "metric": METRIC, "value": value, "unit": "iter/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "ms_per_iter": ms_iter,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "vs_published_rtx5090_opencl": value / world / 2230.0,
            "dtype": "u64 (mod 2^64-2^32+1)", "data": "synthetic (genuine PRP sequence 3^(2^i) mod Mp, dense after the warm-up)",
            "config": {"workload": f"PRP M{HEADLINE_P} (BASELINE configs[1]; ranks > 0: own prime exponent ~1.4e8), "
                                   f"n = 2^23 Marin IBDWT, {iters} square_mul per step, 8 registers",
                       "exponent": HEADLINE_P, "transform_size": n, "iters_per_step": iters,
                       "parallelism": f"replicas x{world} (independent exponents, no data-path collective)",
                       "l2": "working set per iteration (64 MiB register + 128 MiB weight tables) exceeds the 126 MB L2; no flush"},
            "e2e": {"value": e2e_value, "unit": "iter/s", "h2d_bytes_per_step": n * 8, "d2h_bytes_per_step": n * 8,
                    "note": "set_digits from pinned host memory + iters x square_mul + get (canonical digits) per step"},
            "e2e_gerbicz": gerbicz,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": PASSES[dom][0], "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "kernel_ms": {PASSES[i][0]: passes[i] for i in range(len(PASSES))},
                         "note": "integer-pipe bound, not HBM bound: see profiles/ and DESIGN.md"},
            "roofline_square_mul": {"achieved": 48 * n / (ms_iter * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                    "frac": 48 * n / (ms_iter * 1e-3) / 1e9 / peak,
                                    "algorithmic_bytes_per_iter": 48 * n},
        }
        if alu_instr and n == (1 << 23):
            # the unit that actually bounds this path on B200: the integer ALU pipe (one warp instruction per
            # 2 clocks per SM sub-partition); instruction count from the committed ncu capture, time measured here
            sm_hz = ((clocks or {}).get("sm_mhz") or 1965.0) * 1e6
            sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
            peak_alu = sms * 4 * 0.5 * sm_hz
            line["roofline_int_alu_pipe"] = {"achieved": alu_instr / (ms_iter * 1e-3), "peak": peak_alu, "unit": "warp-inst/s",
                                             "frac": alu_instr / (ms_iter * 1e-3) / peak_alu,
                                             "alu_warp_instructions_per_square_mul": alu_instr,
                                             "fmaheavy_pipe_busy_pct_under_ncu": summ.get("fmaheavy_pipe_busy_pct"),
                                             "alu_pipe_busy_pct_under_ncu": summ.get("alu_pipe_busy_pct"),
                                             "note": "every IMAD goes to the FMA-heavy half of the FMA pipe: both integer pipes are busy 54 - 70 % of the "
                                                     "time in the two kernels, math-pipe throttle is their top stall reason (profiles/README.md)",
                                             "source": "profiles/ncu_summary.json (ncu capture) + live timing"}
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"], _ = cpu_baseline(64)
            except Exception as ex:            # the checker failing must not hide the GPU number
                line["cpu_baseline"] = {"value": None, "unit": "iter/s", "cores": os.cpu_count(), "kind": "port",
                                        "sample": f"failed: {ex}"}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)

    L.mb200_host_free(hin); L.mb200_host_free(hout)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
