"""Multi-GPU sharding on real GPUs (run under torchrun, one rank per GPU, NCCL):
  * ECM config 5 (M701, B1 = 6000, B2 = 33333, K = 8 curves, fixed seed): curve c on rank c mod world, results gathered
    and merged in curve order -- must equal the single-process sequential run, curve for curve;
  * P-1 stage 2 (V-trace) with the giant-step range split over the ranks: the product of the gathered accumulators must
    give the factor the single-range run gives;
  * independent PRP exponents, one per rank (the box-throughput configuration), res64 gathered.
Rank 0 also runs the sequential versions and writes gpurun_out/multi_gpu_check_N<world>.json.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/multi_gpu_check.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import prmers_b200 as pb
from prmers_b200 import ecm, modes, multi


def main():
    rank, world = multi.rank_world()
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    factory = lambda p, regs: pb.Engine(p, regs, device=local)
    out = {"world": world}

    # ---- ECM, config 5
    t0 = time.perf_counter()
    sh = ecm.run_ecm_sharded(factory, 701, 6000, 33333, curves=8, seed=123456789)
    out["ecm_seconds_sharded"] = multi.max_over_ranks(time.perf_counter() - t0, device=torch.device("cuda", local))
    key = lambda c: (c["curve"], c["status"], c.get("factor"), c.get("curve_seed"))
    if rank == 0:
        t0 = time.perf_counter()
        seq = ecm.run_ecm(factory, 701, 6000, 33333, curves=8, seed=123456789, continue_after_factor=True)
        out["ecm_seconds_sequential"] = time.perf_counter() - t0
        out["ecm_curves"] = [key(c) for c in sh["curves"]]
        out["ecm_equal_to_sequential"] = [key(c) for c in sh["curves"]] == [key(c) for c in seq["curves"]]
        out["ecm_ranks"] = [c["rank"] for c in sh["curves"]]

    # ---- P-1 stage 2, range split (263: stage-2 factor 321269073670148767; 1362763: the README regression, last giant steps)
    p, b1, b2 = 263, 3527, 16477
    H = modes.run_pm1_stage1(factory(p, 2), p, b1)["x"]
    r = modes.run_pm1_stage2_sharded(factory, p, b1, b2, H, D=210)
    if rank == 0:
        single = modes.run_pm1_stage2_vtrace(factory(p, modes.BABY_BASE + 24), p, b1, b2, H, D=210)
        out["pm1_factor_sharded"] = str(r["factor"]); out["pm1_factor_single"] = str(modes.pm1_stage2_factor(p, [single["acc"]]))
        out["pm1_terms"] = [r["terms"], single["terms"]]
        out["pm1_parts"] = [d["k_range"] for d in r["parts"]]
        out["pm1_equal_to_single"] = r["factor"] == modes.pm1_stage2_factor(p, [single["acc"]]) == 321269073670148767 and r["terms"] == single["terms"]

    # ---- independent exponents, one per rank
    exps = [9941, 9949, 11213, 19937, 21701, 23209, 44497, 86243]
    mine = multi.assign(exps[:max(world, 1)], rank, world)
    res = [{"p": q, "rank": rank, "is_prp": modes.run_prp(factory(q, 8), q)["is_prp"]} for q in mine]
    allres = multi.gather(res)
    if rank == 0:
        out["prp"] = allres
        out["prp_ok"] = all(r["is_prp"] == (r["p"] != 9949) for r in allres) and len(allres) == min(world, len(exps))
        out["ok"] = bool(out["ecm_equal_to_sequential"] and out["pm1_equal_to_single"] and out["prp_ok"])
        os.makedirs("gpurun_out", exist_ok=True)
        with open(f"gpurun_out/multi_gpu_check_N{world}.json", "w") as f:
            json.dump(out, f, indent=1)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
