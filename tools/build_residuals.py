"""The step before the path: build a residual file from a height field on the GPU.

A synthetic DEM of (24 << L)^2 samples -> mipmap levels by point sampling (HeightMipmap::buildMipmapLevel)
-> per level: pl_residual_encode_batch (residual = heights - upsample(parent approximation), int16;
approximation carried to the next level) -> pl_residual_write_file (TIFF/DEFLATE blobs, Lebesgue order,
shared zero blob).  Reports the encode kernel's tiles/s (CUDA events) and the host writer's MB/s, then reads
the file back through pl_residual_decode_batch and checks a sample of tiles.

    python tools/build_residuals.py [--max-level 6] [--out /tmp/DEM.dat]
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "proland-4.0_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import proland_b200 as pl
import resid_synth as rs

MIN_LEVEL, TILE, N = 3, 192, 197


def tile_of(field, ts, tx, ty):
    ix = np.clip(np.arange(-2, ts + 3) + tx * ts, 0, field.shape[1] - 1)
    iy = np.clip(np.arange(-2, ts + 3) + ty * ts, 0, field.shape[0] - 1)
    out = np.zeros((N, N), np.float32)
    out[:ts + 5, :ts + 5] = field[np.ix_(iy, ix)]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-level", type=int, default=6)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    L = a.max_level
    n = 24 << L
    rng = np.random.default_rng(3)
    yy, xx = np.mgrid[0:n + 1, 0:n + 1].astype(np.float32) / n
    base = 1500 * np.sin(6 * xx) * np.cos(5 * yy) + 300 * np.sin(40 * xx + 1) * np.sin(31 * yy) + rng.normal(0, 4, xx.shape).astype(np.float32)
    fields = [np.rint(base).astype(np.int16)]
    for _ in range(L):
        fields.insert(0, fields[0][::2, ::2])
    cnt = lambda l: 1 if l < MIN_LEVEL else 1 << (l - MIN_LEVEL)
    ts_of = lambda l: rs.tile_width(MIN_LEVEL, TILE, l) - 5
    ids = {(l, tx, ty): rs.tile_id(MIN_LEVEL, l, tx, ty) for l in range(L + 1) for ty in range(cnt(l)) for tx in range(cnt(l))}
    nt = len(ids)
    with pl.Context(0) as ctx:
        heights = ctx.pool(pl.POOL_RESID_F32, N, nt)
        approx = ctx.pool(pl.POOL_RESID_F32, N, nt)
        resid = ctx.pool(pl.POOL_RESID_I16, N, nt)
        for (l, tx, ty), tid in ids.items():
            heights.upload(tid, tile_of(fields[l], ts_of(l), tx, ty))
        t0 = tile_of(fields[0], ts_of(0), 0, 0)
        approx.upload(0, t0)
        tiles = {0: np.rint(t0[:ts_of(0) + 5, :ts_of(0) + 5]).astype(np.int16)}
        worst = 0.0
        for timed in (False, True):           # the first pass warms the kernel up (lazy module loading)
          ctx.timing_collect()
          ctx.timing_enable(timed)
          for l in range(1, L + 1):
              keys = [k for k in ids if k[0] == l]
              reqs = np.zeros(len(keys), pl.RESID_ENC_DTYPE)
              for i, (_, tx, ty) in enumerate(keys):
                  parent = (l - 1, tx // 2, ty // 2) if l > MIN_LEVEL else (l - 1, 0, 0)
                  reqs[i] = (ids[(l, tx, ty)], ids[parent], ids[(l, tx, ty)], ids[(l, tx, ty)], ts_of(l), tx, ty, 0)
              mr, me = ctx.residual_encode(heights, approx, resid, reqs)
              worst = max(worst, float(me.max()))
        ms, launches, ntiles = ctx.timing_collect()["residual"]
        ctx.timing_enable(False)
        for (l, tx, ty), tid in ids.items():
            if l:
                w = ts_of(l) + 5
                tiles[tid] = resid.download(tid)[:w, :w]
        out = a.out or os.path.join(tempfile.gettempdir(), "DEM_b200.dat")
        t = time.perf_counter()
        pl.residual_write_file(out, tiles, MIN_LEVEL, L, TILE)
        dt = time.perf_counter() - t
        size = os.path.getsize(out)
        raw = sum(t.size * 2 for t in tiles.values())
        # read back a sample through the device decoder
        import orc
        rd = orc.Resid(open(out, "rb").read())
        sample = sorted(tiles)[::max(1, nt // 64)]
        pool = ctx.pool(pl.POOL_RESID_I16, N, len(sample))
        ctx.residual_decode(pool, [rd.blob(t) for t in sample], [tiles[t].shape[0] for t in sample], list(range(len(sample))))
        for s, tid in enumerate(sample):
            w = tiles[tid].shape[0]
            assert np.array_equal(pool.download(s)[:w, :w], tiles[tid]), tid
        print(json.dumps({"workload": "residual file of a %d^2 DEM: stored levels 0..%d, %d tiles" % (n, L, nt),
                          "encode_tiles_per_s": ntiles / (ms * 1e-3), "encode_ms": ms, "encode_launches": launches,
                          "max_abs_error_m": worst, "file_bytes": size, "compression_ratio": raw / size,
                          "writer_MBps_raw": raw / dt / 1e6, "read_back_tiles_checked": len(sample)}))


if __name__ == "__main__":
    main()
