"""world_size-2 gloo test of the multi-GPU host logic: subtree partition + the counter gather
bench.py does over NCCL (no GPU needed)."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "proland-4.0_b200"))
    import sweep
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    units = sweep.planet_units()
    mine = sweep.units_of_rank(units, rank, world)
    pairs = sweep.pairs_in_units(mine, 10, count_roots=False)
    # what bench.py gathers: per-rank produced pairs, max over ranks of the step time, unit ids
    t = torch.tensor([pairs, 100.0 + rank], dtype=torch.float64)
    gathered = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    ids = torch.tensor([f * 16 + m for f, m in mine], dtype=torch.int64)
    all_ids = [torch.zeros_like(ids) for _ in range(world)]
    dist.all_gather(all_ids, ids)
    tmax = torch.tensor([100.0 + rank], dtype=torch.float64)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((sum(int(g[0]) for g in gathered), float(tmax[0]),
               sorted(int(i) for a in all_ids for i in a)))
    dist.destroy_process_group()


def test_partition_and_gather_world2():
    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    total, tmax, ids = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert total == 8388606 - 6 * 21          # every subtree tile exactly once across the ranks
    assert tmax == 101.0                      # max over ranks
    assert ids == sorted(f * 16 + m for f in range(1, 7) for m in range(16))
