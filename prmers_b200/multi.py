"""Work sharding across the GPUs of one box: replicas only.

A squaring chain is strictly sequential, so one exponent never spans GPUs.  What shards
naturally -- independent exponents, ECM curves, P-1 stage-2 prime ranges -- is dealt round-robin,
unit i to rank i mod world, each rank driving its own engine on its own GPU with no data-path
collective.  The only communication is gathering the (tiny) results at the end, over
torch.distributed (NCCL on the GPUs, gloo in the CPU tests).
"""
import os


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def assign(units, rank, world):
    """Units owned by `rank`: i mod world == rank (the reference's multi-device usage is one process per
    device with its own work list, tests/run_backend_validation_matrix.sh:203-206)."""
    return [u for i, u in enumerate(units) if i % world == rank]


def gather(local_results):
    """All ranks' result lists, concatenated in rank order, on every rank."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(local_results)
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, list(local_results))
    return [r for part in out for r in part]


def max_over_ranks(x, device=None):
    """Max of a float over the ranks (timing is reported as the slowest rank)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
