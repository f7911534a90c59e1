"""Kernel time of pl_residual_decode_batch (inflate + store kernels, the library's CUDA events) on config 3's
residual tiles: 4096 TIFF/DEFLATE blobs of 197 x 197 int16 (64 distinct synthetic SRTM-shaped tiles, zlib level 6).
    python tools/inflate_time.py [reps]          (PL_LIB selects a build variant)"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "proland-4.0_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import proland_b200 as pl
import resid_synth as rs

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
rng = np.random.default_rng(20240612)
distinct = [rs.fractal_tile(rng, 197, 40) for _ in range(64)]
blobs64 = [rs.tiff_blob(t, 6) for t in distinct]
nres = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
blobs = [blobs64[i % 64] for i in range(nres)]
in_bytes = sum(len(b) for b in blobs)
with pl.Context(0) as ctx:
    pool = ctx.pool(pl.POOL_RESID_I16, 197, nres)
    ctx.residual_decode(pool, blobs, [197] * nres, list(range(nres)), scale=1.0)
    ctx.timing_collect()
    ctx.timing_enable(True)
    for _ in range(reps):
        ctx.residual_decode(pool, blobs, [197] * nres, list(range(nres)), scale=1.0)
    ms = ctx.timing_collect()["residual"][0] / reps
    ctx.timing_enable(False)
    for k in (0, 5, 63, nres - 1):
        assert np.array_equal(pool.download(k)[:197, :197].astype(np.int16), distinct[k % 64]), "decode differs from the source tile"
    print(json.dumps({"lib": os.environ.get("PL_LIB", "default"), "tiles_per_s": nres / (ms * 1e-3), "ms_per_batch": ms,
                      "decoded_GBps": nres * 197 * 197 * 2 / (ms * 1e-3) / 1e9,
                      "compression_ratio": nres * 197 * 197 * 2 / in_bytes}))
