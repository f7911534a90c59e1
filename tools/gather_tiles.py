"""Gathering finished tiles and statistics over NCCL / NVLink (north star: "NCCL is used only to gather finished
tiles or statistics"; SURVEY 8e: measured as a separate number, not part of the production path).

Every rank produces its contiguous Morton range of each level of one planet face (levels below the first
level that splits evenly are replicated) into a pool laid out for the WHOLE quadtree, so that the tiles of a level
form one slab per pool and a rank's share is one contiguous piece of it.  Then, for the deepest level:

  stats      all_gather of the per-tile (zmin, zmax) pairs (8 B per tile)            TileSamplerZ's consumers
  normals    in-place all_gather of the RG8 normal slab (18 848 B per tile)           a renderer on every GPU
  elevations in-place all_gather of the elevation slab (126 048 B per tile)

timed with CUDA events on the launching stream (max over ranks), reported as bytes received per rank and second.
Afterwards rank 0 produces the whole level by itself and compares: the gathered slabs must be bit-identical.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/gather_tiles.py [--level 7]
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "proland-4.0_b200"))
sys.path.insert(0, ROOT)
import proland_b200 as pl

PLANET = [-3250, -1590, -1125, -795, -561, -397, -140, -100, 15, 8, 5, 2.5, 1.5, 1, 0.5, 0.25, 0.1, 0.05]


def split_level(world):
    """the first level whose tile count splits evenly over the ranks"""
    l = 0
    while 4 ** l < world or 4 ** l % world:
        l += 1
    return l


def rank_range(level, rank, world):
    """(morton0, n) of the tiles of `level` a rank produces; whole levels above the split level"""
    if level < split_level(world):
        return 0, 4 ** level
    n = 4 ** level // world
    return rank * n, n


def share_fd(dist, rank, world, fd, key):
    """hands rank 0's file descriptor to every other rank of the box: SCM_RIGHTS over an abstract Unix socket (the C ABI
    leaves the inter-process transport to the caller).  -> the local descriptor.  Every rank reaches the barrier whatever
    happens; a failure surfaces as OSError on the ranks it concerns (sockets time out instead of waiting for ever)."""
    import socket
    name = "\0proland-b200-mc-%s-%s" % (os.environ.get("MASTER_PORT", "0"), key)
    srv, err = None, None
    if rank == 0:
        try:
            srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            srv.bind(name)
            srv.listen(world)
            srv.settimeout(30.0)
        except OSError as e:
            err = e
    dist.barrier()                       # the socket is listening (or rank 0 failed: the peers' connect is refused)
    if rank == 0:
        if err is not None:
            raise err
        try:
            for _ in range(world - 1):
                conn, _ = srv.accept()
                conn.settimeout(30.0)
                socket.send_fds(conn, [b"fd"], [fd])
                conn.recv(1)             # the peer has the descriptor
                conn.close()
        finally:
            srv.close()
        return fd
    c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    c.settimeout(30.0)
    try:
        c.connect(name)
        _, fds, _, _ = socket.recv_fds(c, 16, 1)
        c.send(b"k")
    finally:
        c.close()
    if not fds:
        raise OSError("no file descriptor received")
    return fds[0]


def multicast_group(dist, pool, rank, world, key="norm"):
    """binds a shared pool (ctx.pool(..., shared=True)) of every rank to one NVLink multicast object.  Every stage ends
    with an agreement over the ranks (all_reduce MIN of a success flag, which is also the barrier the C ABI asks for): if a
    stage fails anywhere, ALL ranks give the group up together instead of waiting for each other.  -> True when bound"""
    import torch

    def agreed(ok):
        t = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return float(t[0]) == 1.0

    fd, ok = -1, True
    try:
        if rank == 0:
            fd = pool.mc_create(world)
    except pl.PlError:
        ok = False
    if not agreed(ok):
        return False
    try:
        fd = share_fd(dist, rank, world, fd, key)
        if rank != 0:
            pool.mc_import(fd, world)
        os.close(fd)
        pool.mc_add_device()
    except (pl.PlError, OSError):
        ok = False
    if not agreed(ok):                   # every device is added: binding may start
        return False
    try:
        pool.mc_bind()
    except pl.PlError:
        ok = False
    return agreed(ok)


class _DeviceBytes:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def gather_record(ctx, torch, dist, stream, rank, world, level=7, reps=5):
    """the gather measurements on an existing context / stream (bench.py calls this at N > 1); -> the record on rank 0,
    None elsewhere.  Collective: every rank of the group must call it."""
    line = None
    L = level
    off = [(4 ** l - 1) // 3 for l in range(L + 2)]
    if True:
        elev = ctx.pool(pl.POOL_ELEV, 101, off[L + 1])
        norm = ctx.pool(pl.POOL_NORM2, 97, off[L + 1])
        ctx.noise_init(101)
        sc = pl.sweep_scene(noise_amp=PLANET, face=3, root_quad_size=12720000.0, sphere=1, want_stats=1)

        def produce(r, w, norm=norm):
            for l in range(L + 1):
                m0, n = rank_range(l, r, w)
                ctx.produce_range(sc, elev, norm, l, m0, n, off[l] + m0, off[l - 1] + (m0 >> 2) if l else 0, m0 >> 2)

        def slab(pool):
            base = pl.lib().pl_pool_device_ptr(pool.h) + off[L] * pool.slot_bytes
            return torch.as_tensor(_DeviceBytes(base, 4 ** L * pool.slot_bytes), device="cuda")

        with torch.cuda.stream(stream):
            produce(rank, world)
            ctx.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(reps):
                produce(rank, world)
            e1.record(stream)
            torch.cuda.synchronize()
            prod_ms = e0.elapsed_time(e1) / reps

            m0, n = rank_range(L, rank, world)
            results = {}
            # statistics: what TileSamplerZ reads back, 8 bytes per tile
            mine = torch.from_numpy(ctx.elev_stats_range(elev, off[L] + m0, n)).cuda()
            allst = torch.empty((4 ** L, 2), dtype=torch.float32, device="cuda")
            for name, out, inp in (("stats", allst, mine),):
                dist.all_gather_into_tensor(out, inp)
                torch.cuda.synchronize()
                e0.record(stream)
                for _ in range(reps):
                    dist.all_gather_into_tensor(out, inp)
                e1.record(stream)
                torch.cuda.synchronize()
                results[name] = e0.elapsed_time(e1) / reps
            for name, pool in (("normals", norm), ("elevations", elev)):
                full = slab(pool)
                part = full[m0 * pool.slot_bytes:(m0 + n) * pool.slot_bytes]
                dist.all_gather_into_tensor(full, part)          # in place: a rank's share is already where it belongs
                torch.cuda.synchronize()
                e0.record(stream)
                for _ in range(reps):
                    dist.all_gather_into_tensor(full, part)
                e1.record(stream)
                torch.cuda.synchronize()
                results[name] = e0.elapsed_time(e1) / reps
            t = torch.tensor([prod_ms] + [results[k] for k in ("stats", "normals", "elevations")], dtype=torch.float64,
                             device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            # parity of the gather: every rank now holds the whole level; rank 0 produces it alone and compares
            got_n, got_e = slab(norm).clone(), slab(elev).clone()
            sums = torch.stack([got_n.sum(dtype=torch.int64), got_e.sum(dtype=torch.int64)])
            allsums = [torch.zeros_like(sums) for _ in range(world)]
            dist.all_gather(allsums, sums)
            identical = all(bool(torch.equal(s, allsums[0])) for s in allsums)
            if rank == 0:
                for r in range(world):
                    produce(r, world)
                ctx.sync()
                identical = identical and bool(torch.equal(slab(norm), got_n)) and bool(torch.equal(slab(elev), got_e))
                st = ctx.elev_stats_range(elev, off[L], 4 ** L)
                identical = identical and np.array_equal(st, allst.cpu().numpy())

            # ---- the same gather WITHOUT a collective: the fused kernel stores every finished normal tile into the
            # peers' pools over NVLink while it produces (pl_pool_attach_peers / pl_pool_push_to_peers)
            push = None
            if world > 1:
                mine_h = torch.from_numpy(norm.export()).cuda()
                handles = [torch.zeros_like(mine_h) for _ in range(world)]
                dist.all_gather(handles, mine_h)
                norm.attach_peers(torch.stack(handles).cpu().numpy(), rank)
                slab(norm).zero_()
                torch.cuda.synchronize()
                dist.barrier()
                norm.push_to_peers(True)
                produce(rank, world)
                ctx.sync()
                dist.barrier()
                torch.cuda.synchronize()
                pushed_ok = bool(torch.equal(slab(norm), got_n))        # every rank holds the whole level again
                e0.record(stream)
                for _ in range(reps):
                    produce(rank, world)
                e1.record(stream)
                torch.cuda.synchronize()
                dist.barrier()
                norm.push_to_peers(False)
                tp = torch.tensor([e0.elapsed_time(e1) / reps, 0.0 if pushed_ok else 1.0], dtype=torch.float64, device="cuda")
                dist.all_reduce(tp, op=dist.ReduceOp.MAX)
                push = (float(tp[0]), float(tp[1]) == 0.0)

            # ---- and through ONE store per 16 bytes: a pool on the VMM allocator, bound on every rank to one NVLink
            # multicast object; the switch replicates the store into every GPU's pool (pl_pool_mc_*, pl_multicast.cu)
            mcast = None
            if world > 1 and os.environ.get("PL_NO_MULTICAST") is None:
                try:
                    normm = ctx.pool(pl.POOL_NORM2, 97, off[L + 1], shared=True)
                    have = 1.0
                except pl.PlError:
                    normm, have = None, 0.0
                hv = torch.tensor([have], dtype=torch.float64, device="cuda")
                dist.all_reduce(hv, op=dist.ReduceOp.MIN)
                if float(hv[0]) == 1.0 and multicast_group(dist, normm, rank, world):
                    slab(normm).zero_()
                    torch.cuda.synchronize()
                    dist.barrier()
                    normm.push_to_peers(2)
                    produce(rank, world, normm)
                    ctx.sync()
                    dist.barrier()
                    torch.cuda.synchronize()
                    mc_ok = bool(torch.equal(slab(normm), got_n))
                    e0.record(stream)
                    for _ in range(reps):
                        produce(rank, world, normm)
                    e1.record(stream)
                    torch.cuda.synchronize()
                    dist.barrier()
                    normm.push_to_peers(False)
                    tm = torch.tensor([e0.elapsed_time(e1) / reps, 0.0 if mc_ok else 1.0], dtype=torch.float64, device="cuda")
                    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
                    mcast = (float(tm[0]), float(tm[1]) == 0.0)
                    dist.barrier()
                if normm is not None:
                    normm.close()
        if rank == 0:
            tiles = 4 ** L
            recv = lambda pool_bytes: (world - 1) / world * tiles * pool_bytes
            line = {"workload": "gather of the %d finished tiles of level %d of one planet face, %d ranks, "
                                "in-place NCCL all_gather" % (tiles, L, world),
                    "n_gpus": world, "production_ms": float(t[0]),
                    "production_pairs_per_s": sum(rank_range(l, 0, world)[1] for l in range(L + 1)) * world / (float(t[0]) * 1e-3),
                    "identical_to_single_gpu": bool(identical)}
            for i, (k, b) in enumerate((("stats", 8), ("normals", norm.slot_bytes), ("elevations", elev.slot_bytes))):
                ms = float(t[1 + i])
                line[k] = {"ms": ms, "bytes_received_per_rank": recv(b), "GBps_per_rank": recv(b) / (ms * 1e-3) / 1e9}
            if push:
                per_rank_tiles = sum(rank_range(l, 0, world)[1] for l in range(L + 1))
                sent = (world - 1) * per_rank_tiles * 97 * 97 * 2
                line["push_from_the_kernel"] = {
                    "what": "production of levels 0..%d with every finished RG8 normal tile also stored into the peers' "
                            "pools by the fused kernel (no collective)" % L,
                    "production_ms": push[0], "extra_ms_vs_plain_production": push[0] - float(t[0]),
                    "all_gather_of_the_normals_ms": float(t[2]), "bytes_sent_per_rank": sent,
                    "identical_on_every_rank": push[1]}
            if mcast:
                line["multicast_push_from_the_kernel"] = {
                    "what": "the same, one store per 16 bytes through an NVLink multicast mapping of the normal pools (cuMulticast: "
                            "the switch replicates it into every GPU's pool)",
                    "production_ms": mcast[0], "extra_ms_vs_plain_production": mcast[0] - float(t[0]),
                    "bytes_sent_per_rank": sum(rank_range(l, 0, world)[1] for l in range(L + 1)) * 97 * 97 * 2,
                    "identical_on_every_rank": mcast[1]}
    dist.barrier()              # nobody still reads a peer's pool
    torch.cuda.synchronize()
    norm.close()
    elev.close()
    return line


def main():
    import torch
    import torch.distributed as dist
    ap = argparse.ArgumentParser()
    ap.add_argument("--level", type=int, default=7)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)                   # NCCL's version banner goes to stderr: stdout is the one JSON line
    try:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.all_reduce(torch.zeros(1, device="cuda"))
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        ctypes.CDLL(None).fflush(None)          # the banner sits in C stdio's buffer when stdout is a pipe
        os.dup2(saved, 1)
        os.close(saved)
    with pl.Context(local) as ctx:
        stream = torch.cuda.Stream()
        ctx.set_stream(stream.cuda_stream)
        line = gather_record(ctx, torch, dist, stream, rank, world, a.level, a.reps)
        if rank == 0:
            print(json.dumps(line), flush=True)
    dist.destroy_process_group()



if __name__ == "__main__":
    main()
