"""Config 3 (BASELINE.json, demo-earth-srtm shape) throughput on one B200.

  1. residual decode: 4096 TIFF/DEFLATE blobs of 197 x 197 int16 residuals (64 distinct synthetic
     SRTM-shaped tiles, zlib level 6, cycled) -> the int16 / float residual pool
     (pl_residual_decode_batch: the warp-per-tile inflate kernel + the pitched store kernel)
  2. elevation + normal pairs WITH residuals: one level-7 batch of a sphere face (16 384 tiles,
     flip, NEAREST elevation storage, one 197-wide residual tile per 2 x 2 tiles) through
     pl_pair_batch -- the fused kernel's RESID variants -- on the int16 pool (residuals consumed as
     stored: 212 500 algorithmic bytes per pair) and on the float pool (232 902 B per pair).

Timings are the library's CUDA events around the launches.  One JSON line per measurement.

    python tools/srtm_sweep.py [--reps 5]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "proland-4.0_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import proland_b200 as pl
import resid_synth as rs
import bench

AMP = [0] * 7 + [5, 2.5, 1]          # earth-srtm.xml: no noise down to the residual levels, then a little
SIZE = 12720000.0
LEVEL = 7


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    peak, peak_kind = bench.peaks()
    rng = np.random.default_rng(20240612)
    distinct = [rs.fractal_tile(rng, 197, 40) for _ in range(64)]
    blobs64 = [rs.tiff_blob(t, 6) for t in distinct]
    nres = 4 ** (LEVEL - 1)
    blobs = [blobs64[i % 64] for i in range(nres)]
    in_bytes = sum(len(b) for b in blobs)
    with pl.Context(0) as ctx:
        ctx.noise_init(101)
        pools = {}
        # ---------------------------------------------------------------- 1. residual decode
        for kind, name, esz in ((pl.POOL_RESID_I16, "int16", 2), (pl.POOL_RESID_F32, "float", 4)):
            pool = ctx.pool(kind, 197, nres)
            ctx.residual_decode(pool, blobs, [197] * nres, list(range(nres)), scale=1.0)     # warm-up
            ctx.timing_collect()
            ctx.timing_enable(True)
            for _ in range(args.reps):
                ctx.residual_decode(pool, blobs, [197] * nres, list(range(nres)), scale=1.0)
            ms = ctx.timing_collect()["residual"][0] / args.reps
            ctx.timing_enable(False)
            got = pool.download(5)[:197, :197]
            assert np.array_equal(got.astype(np.int16), distinct[5]), "decode differs from the source tile"
            out_bytes = nres * 197 * 197 * esz
            print(json.dumps({"workload": "config 3 residual decode, %d blobs of 197x197 int16 -> %s pool" % (nres, name),
                              "tiles_per_s": nres / (ms * 1e-3), "ms_per_batch": ms,
                              "compressed_MBps": in_bytes / (ms * 1e-3) / 1e6,
                              "decoded_GBps": out_bytes / (ms * 1e-3) / 1e9,
                              "compression_ratio": nres * 197 * 197 * 2 / in_bytes}), flush=True)
            pools[name] = pool

        # ---------------------------------------------------------------- 2. pairs with residuals
        off = [sum(4 ** k for k in range(l)) for l in range(LEVEL + 2)]
        elev = ctx.pool(pl.POOL_ELEV, 101, off[LEVEL + 1])
        norm = ctx.pool(pl.POOL_NORM2, 97, off[LEVEL + 1])
        sc = pl.sweep_scene(noise_amp=AMP, face=2, root_quad_size=SIZE, sphere=1, flip=1,
                            elev_filter=pl.FILTER_NEAREST, want_stats=1)
        for l in range(LEVEL):      # ancestors: fractal only
            ctx.produce_range(sc, elev, norm, l, 0, 4 ** l, off[l], off[l - 1] if l else 0, 0)
        n = 4 ** LEVEL
        e, q = pl.make_requests_range(sc, LEVEL, 0, n, off[LEVEL], off[LEVEL - 1], 0)
        # one residual tile per 2 x 2 tiles: window origin (tx % 2, ty % 2) * 96 inside the 197-wide tile
        tx, ty = e["tx"], e["ty"]
        e["resid_slot"] = (tx // 2) + (ty // 2) * (1 << (LEVEL - 1))
        e["rx"] = (tx % 2) * 96
        e["ry"] = (ty % 2) * 96
        for name, per_pair in (("int16", 212500), ("float", 232902)):
            es = pl.elev_scene(101, 24, 1, 1, 0, 1, resid_scale=1.0)
            ctx.pair_batch(es, sc.norm, elev, norm, e, q, resid=pools[name])      # warm-up
            ctx.timing_collect()
            ctx.timing_enable(True)
            for _ in range(args.reps):
                ctx.pair_batch(es, sc.norm, elev, norm, e, q, resid=pools[name])
            ms = ctx.timing_collect()["pair"][0] / args.reps
            ctx.timing_enable(False)
            gbs = per_pair * n / (ms * 1e-3) / 1e9
            print(json.dumps({"workload": "config 3 elevation+normal with residuals (%s pool), %d level-%d tiles, sphere, flip, NEAREST" % (name, n, LEVEL),
                              "pairs_per_s": n / (ms * 1e-3), "ms_per_batch": ms,
                              "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s",
                                           "frac": gbs / peak, "bytes_per_pair": per_pair, "peak_kind": peak_kind}}), flush=True)


if __name__ == "__main__":
    main()
