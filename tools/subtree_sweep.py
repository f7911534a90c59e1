"""Config 4 (BASELINE.json): full-subtree batch sweep -- all level-14 descendants of one tile at level
14 - d (d = 6 .. 10: 4 096 .. 1 048 576 leaf tiles, plus a third as many ancestors), produced level by
level with pl_produce_range and partitioned by subtree (sweep.SubtreeSweep).  Flat face 0 with the
demo-fractalterrain noise amplitudes (zero past level 11), flat RG8 normals.

Reports tiles/s (device time between CUDA events) and an order-independent fingerprint of the result
(the per-tile (zmin, zmax) statistics of every leaf tile), which must not depend on how the sweep was
partitioned.

    python tools/subtree_sweep.py [--d 6 7 8 9 10] [--world 1] [--unit-depth 7]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "proland-4.0_b200"))

FRACTAL = [-140, -100, -15, -8, 5, 2.5, 1.5, 1, 0.5, 0.25, 0.1, 0.05]
LEAF_LEVEL = 14


def run_sweep(pl, ctx, d, rank=0, world=1, unit_depth=7, root=None, scene_kw=None, keep=None, levels=False, pools=None):
    """Produces rank `rank`'s share; returns (tiles produced, fingerprint dict).  keep: optional dict
    filled with {(level, tx, ty): (elev, norm)} for the leaf tiles whose Morton index is in keep['want'].
    levels: the chain + top tree, and every unit, as ONE launch each (pl_produce_levels: the level-to-level dependency
    is resolved inside the kernel) instead of one pl_produce_range per level.
    pools: (elev, norm) of at least plan.capacity slots to produce into (kept open); default: created and closed here."""
    import sweep
    root_level = LEAF_LEVEL - d
    tx, ty = root if root is not None else ((1 << root_level) // 3, (1 << root_level) // 5)
    plan = sweep.SubtreeSweep(root_level, tx, ty, d, unit_depth)
    kw = dict(noise_amp=FRACTAL, face=0, root_quad_size=100000.0, sphere=0, want_stats=1)
    kw.update(scene_kw or {})
    sc = pl.sweep_scene(**kw)
    if pools is None:
        elev = ctx.pool(pl.POOL_ELEV, 101, plan.capacity)
        norm = ctx.pool(pl.POOL_NORM2, 97, plan.capacity)
        ctx.noise_init(101)
    else:
        elev, norm = pools
    produced = 0
    carry = []       # levels mode, k == 0: the chain and the single unit are one consecutive run of levels
    if levels:
        carry = list(plan.prologue())
        produced += sum(b[2] for b in carry)
        if plan.k > 0:
            ctx.produce_levels(sc, elev, norm, carry)
            carry = []
    else:
        for level, m0, n, s0, p0, pm0 in plan.prologue():
            ctx.produce_range(sc, elev, norm, level, m0, n, s0, p0, pm0)
            produced += n
    slot0, nleaf = plan.leaf_region()
    fp_sum, fp_lo, fp_hi, fp_xor = 0.0, np.inf, -np.inf, np.uint32(0)
    def fold(st):
        nonlocal fp_sum, fp_lo, fp_hi, fp_xor
        fp_sum += float(st.astype(np.float64).sum())
        fp_lo, fp_hi = min(fp_lo, float(st[:, 0].min())), max(fp_hi, float(st[:, 1].max()))
        fp_xor ^= np.bitwise_xor.reduce(st.view(np.uint32).ravel())

    pending = []      # TileSamplerZ's readback, 8 bytes per leaf tile: enqueued behind the unit's kernels,
    for unit in plan.units_of_rank(rank, world):          # collected two units later (ReadbackManager)
        if levels:
            ub = list(plan.unit_batches(unit))
            ctx.produce_levels(sc, elev, norm, carry + ub)
            carry = []
            produced += sum(b[2] for b in ub)
        else:
            for level, m0, n, s0, p0, pm0 in plan.unit_batches(unit):
                ctx.produce_range(sc, elev, norm, level, m0, n, s0, p0, pm0)
                produced += n
        if len(pending) == 2:
            fold(ctx.elev_stats_readback_end(pending.pop(0)))
        pending.append(ctx.elev_stats_readback_begin(elev, slot0, nleaf))
        if keep is not None:
            leaf_m0 = ((plan.root_morton << (2 * plan.k)) | unit) << (2 * (d - plan.k))
            for m in keep["want"]:
                if leaf_m0 <= m < leaf_m0 + nleaf:
                    s = slot0 + (m - leaf_m0)
                    keep[m] = (elev.download(s), norm.download(s))
    for tk in pending:
        fold(ctx.elev_stats_readback_end(tk))
    ctx.sync()
    if pools is None:
        elev.close()
        norm.close()
    return produced, dict(sum=fp_sum, lo=fp_lo, hi=fp_hi, xor=int(fp_xor)), plan


if __name__ == "__main__":
    import torch
    import proland_b200 as pl
    ap = argparse.ArgumentParser()
    ap.add_argument("--d", type=int, nargs="+", default=[6, 7, 8, 9, 10])
    ap.add_argument("--world", type=int, default=1, help="partition into this many ranks (run one after the other here)")
    ap.add_argument("--unit-depth", type=int, default=7)
    a = ap.parse_args()
    with pl.Context(0) as ctx:
        stream = torch.cuda.Stream()
        ctx.set_stream(stream.cuda_stream)
        for d in a.d:
            run_sweep(pl, ctx, min(d, 6))                       # warm-up
            fps, total, ms = [], 0, 0.0
            for rank in range(a.world):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                n, fp, plan = run_sweep(pl, ctx, d, rank, a.world, a.unit_depth)
                e1.record(stream)
                torch.cuda.synchronize()
                ms = max(ms, e0.elapsed_time(e1))               # ranks would run concurrently: max over ranks
                total += n
                fps.append(fp)
            total -= (a.world - 1) * plan.replicated_tiles()    # the replicated ancestors count once
            print(json.dumps({"workload": "config 4 subtree sweep, d=%d: %d leaf tiles of level 14 under (%d,%d,%d)"
                              % (d, 4 ** d, plan.root_level, plan.tx, plan.ty), "tiles": total, "world": a.world,
                              "ms_max_over_ranks": ms, "pairs_per_s": total / (ms * 1e-3),
                              "fingerprint": {"sum": sum(f["sum"] for f in fps), "lo": min(f["lo"] for f in fps),
                                              "hi": max(f["hi"] for f in fps), "xor": int(np.bitwise_xor.reduce([f["xor"] for f in fps]))}}))
