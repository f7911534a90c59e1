"""Config 1 (BASELINE.json): demo-fractalterrain -- the flat fractal terrain of
src/terrain/examples/terrain1 / fractalterrain.xml: face 0, no residuals,
noise="-140,-100,-15,-8,5,2.5,1.5,1,0.5,0.25,0.1,0.05", LINEAR elevation storage, flat RG8 normals,
the full quadtree of levels 0..8 (87 381 elevation+normal pairs), breadth first, resident pool.

Reports pairs/s (CUDA events around the whole sweep and per launch) and the checksum of checksums
(sum over tiles of per-tile zmin + zmax) which tests/test_gpu_parity.py pins against the oracle on
levels 0..4 of the same scene.

    python tools/fractalterrain.py [--max-level 8] [--reps 5]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "proland-4.0_b200"))
sys.path.insert(0, ROOT)
import proland_b200 as pl
import bench

FRACTAL = [-140, -100, -15, -8, 5, 2.5, 1.5, 1, 0.5, 0.25, 0.1, 0.05]


def run(pl, ctx, torch, stream, max_level=8, reps=5, arith=0):
    """One record: the sweep's pairs/s (CUDA events around `reps` sweeps on `stream`, the context's stream) and the
    fused kernel's own roofline (the library's per-launch events)."""
    peak, peak_kind = bench.peaks()
    off = [sum(4 ** k for k in range(l)) for l in range(max_level + 2)]
    total = off[max_level + 1]
    elev = ctx.pool(pl.POOL_ELEV, 101, total)
    norm = ctx.pool(pl.POOL_NORM2, 97, total)
    ctx.noise_init(101)
    sc = pl.sweep_scene(noise_amp=FRACTAL, face=0, root_quad_size=100000.0, sphere=0,
                        elev_filter=pl.FILTER_LINEAR, want_stats=1, arith=arith)

    def sweep():
        for l in range(max_level + 1):
            ctx.produce_range(sc, elev, norm, l, 0, 4 ** l, off[l], off[l - 1] if l else 0, 0)
    sweep()
    ctx.sync()
    ctx.timing_collect()
    ctx.timing_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        sweep()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    kt = ctx.timing_collect()
    ctx.timing_enable(False)
    st = ctx.elev_stats_range(elev, 0, total).astype(np.float64)
    pair_ms, launches, tiles = kt["pair"]
    gbs = bench.PAIR_BYTES * tiles / (pair_ms * 1e-3) / 1e9
    norm.close()
    elev.close()
    return {"workload": "config 1 demo-fractalterrain: flat face, levels 0..%d, %d pairs per sweep" % (max_level, total),
            "arith": "fast" if arith else "exact",
            "pairs_per_s": total / (ms * 1e-3), "ms_per_sweep": ms, "launches_per_sweep": 2 * (launches // reps),
            "roofline": {"bound": "hbm", "kernel": "pair", "achieved": gbs, "peak": peak, "unit": "GB/s",
                         "frac": gbs / peak, "bytes_per_pair": bench.PAIR_BYTES, "peak_kind": peak_kind,
                         "kernel_ms_per_sweep": pair_ms / reps, "share_of_sweep": pair_ms / reps / ms,
                         "traffic_per_pair": bench.TRAFFIC.get("pair_flat")},
            "fingerprint": {"sum": float(st.sum()), "zmin": float(st[:, 0].min()), "zmax": float(st[:, 1].max())}}


def main():
    import torch
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-level", type=int, default=8)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--fast", action="store_true", help="PL_ARITH_FAST normals")
    a = ap.parse_args()
    with pl.Context(0) as ctx:
        stream = torch.cuda.Stream()
        ctx.set_stream(stream.cuda_stream)
        print(json.dumps(run(pl, ctx, torch, stream, a.max_level, a.reps, int(a.fast))))


if __name__ == "__main__":
    main()
