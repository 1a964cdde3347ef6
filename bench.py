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
PASSES = [("k_mid (contiguous pass: forward, pointwise square, inverse)", 1, 16),
          ("k_strided<1> (strided inverse + unweight/carry/reweight + next strided forward)", 3, 32)]


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
                      f"oracle/marin_oracle.c with OpenMP; the reference's OpenCL path needs PoCL + OpenCL/GMP "
                      f"headers, absent from this image"}, o


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path = the oracle port (the stock
    binary cannot be built here: no CL/cl.h, no gmp.h, no PoCL)."""
    if rank != 0:
        return
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count())      # torchrun pins it to 1; the reference arm uses every core
    from oracle import marin_oracle as mo
    sample = 2
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
                             "sample": f"{sample} square_mul per step on all host threads (OpenMP), oracle/marin_oracle.c"},
            "e2e": {"value": v, "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--iters", type=int, default=ITERS, help="square_mul per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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

    if rank == 0:
        peak, peak_src = measured_peak()
        # per-kernel durations, live, with CUDA events on the launching stream
        passes = []
        for _, which, _ in PASSES:
            cnt = 100
            passes.append(eng.bench_pass(1, which, cnt) / cnt)            # ms per launch
        dom = max(range(len(PASSES)), key=lambda i: passes[i])
        alg_bytes = PASSES[dom][2] * n                                    # 8n read + 8n write per pass of the 3-pass count
        achieved = alg_bytes / (passes[dom] * 1e-3) / 1e9
        traffic, alu_instr = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_summary.json")) as f:
                summ = json.load(f)
            traffic = summ["dram_bytes_per_launch"].get(["k_mid", "k_strided<1>+fwd"][dom])
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
                                             "source": "profiles/ncu_summary.json (ncu capture r1n) + live timing"}
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
