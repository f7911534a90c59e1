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


def test_gather_partition_covers_every_level_once():
    """tools/gather_tiles.py: contiguous Morton ranges per rank -- above the split level every tile of a level belongs to
    exactly one rank, a rank's range of level l + 1 is the children of its range of level l (parents stay local), and a
    rank's share is one contiguous piece of the level's slab (what makes the all_gather in place)"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gather_tiles", os.path.join(ROOT, "tools", "gather_tiles.py"))
    gt = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gt)
    for world in (1, 2, 4, 8):
        ls = gt.split_level(world)
        assert 4 ** ls % world == 0 and (ls == 0 or 4 ** (ls - 1) % world or 4 ** (ls - 1) < world)
        for level in range(ls, 8):
            ranges = [gt.rank_range(level, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and sum(n for _, n in ranges) == 4 ** level
            for (m0, n), (m1, _) in zip(ranges, ranges[1:]):
                assert m0 + n == m1
            for r in range(world):
                m0, n = ranges[r]
                c0, cn = gt.rank_range(level + 1, r, world)
                assert (c0, cn) == (4 * m0, 4 * n)
        for level in range(ls):
            assert all(gt.rank_range(level, r, world) == (0, 4 ** level) for r in range(world))


def _fd_worker(rank, world, port, q):
    import importlib.util
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    spec = importlib.util.spec_from_file_location("gather_tiles", os.path.join(ROOT, "tools", "gather_tiles.py"))
    gt = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gt)
    # rank 0 owns a pipe; its write end travels to the other ranks as a file descriptor (what the multicast object's
    # descriptor does, tools/gather_tiles.py::share_fd)
    r_end, w_end = os.pipe() if rank == 0 else (-1, -1)
    fd = gt.share_fd(dist, rank, world, w_end, "test")
    os.write(fd, bytes([65 + rank]))
    dist.barrier()
    if rank == 0:
        q.put(sorted(os.read(r_end, 16)))
    dist.destroy_process_group()


def test_share_fd_hands_a_descriptor_to_every_rank():
    """the transport the multicast push leaves to its caller: SCM_RIGHTS over an abstract Unix socket between the ranks of a
    box (world 3, gloo): every rank ends up with a descriptor of rank 0's pipe and can write into it"""
    world = 3
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_fd_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == [65, 66, 67]
