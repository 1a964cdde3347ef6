"""CPU tier: the N > 1 host logic (sharding + result gathering) over gloo, world size 2."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from prmers_b200 import multi


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    units = [136279841, 139123459, 139246931, 139370393, 139493863]
    mine = multi.assign(units, rank, world)
    local = [{"p": p, "rank": rank, "res64": f"{p * 2654435761 % (1 << 64):016X}"} for p in mine]
    everything = multi.gather(local)
    slowest = multi.max_over_ranks(10.0 + rank)
    q.put((rank, mine, everything, slowest))
    dist.barrier()
    dist.destroy_process_group()


def test_round_robin_sharding_and_gather_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    units = [136279841, 139123459, 139246931, 139370393, 139493863]
    assert got[0][1] == units[0::2] and got[1][1] == units[1::2]
    for rank, _, everything, slowest in got:
        assert sorted(r["p"] for r in everything) == sorted(units)          # nothing lost, nothing twice
        assert [r["rank"] for r in everything] == [0, 0, 0, 1, 1]            # rank order
        assert slowest == 11.0


def _ecm_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.int_engine import IntEngine
    from prmers_b200 import ecm
    r = ecm.run_ecm_sharded(lambda p, regs: IntEngine(p, regs), 701, 300, 900, curves=5, seed=123456789)
    q.put((rank, r["status"], r["factor"], [(c["curve"], c["rank"], c["status"], c["factor"]) for c in r["curves"]]))
    dist.barrier()
    dist.destroy_process_group()


def test_ecm_curves_sharded_world2_equals_sequential():
    """Curve c on rank c mod 2; the merged result is what one process running the curves in order reports."""
    from oracle.int_engine import IntEngine
    from prmers_b200 import ecm
    seq = ecm.run_ecm(lambda p, regs: IntEngine(p, regs), 701, 300, 900, curves=5, seed=123456789, continue_after_factor=True)
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ecm_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, status, factor, curves in got:
        assert (status, factor) == (seq["status"], seq["factor"])
        assert [c[1] for c in curves] == [0, 1, 0, 1, 0]
        assert [(c[0], c[2], c[3]) for c in curves] == [(c["curve"], c["status"], c["factor"]) for c in seq["curves"]]


def test_single_process_paths():
    assert multi.assign(list(range(7)), 1, 3) == [1, 4]
    assert multi.gather([1, 2]) == [1, 2]
    assert multi.max_over_ranks(3.5) == 3.5
    assert multi.rank_world() == (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")))
