"""Ortho path: terrain3/helloworld.xml's orthoProducer (hsv noise, rnoise 60,150,20,
cnoise 70,80,100, amplitudes 255, 196-texel tiles) and a plain-noise variant: the full quadtree of levels
0..L of one face, breadth first, resident pool, through pl_ortho_batch with HOST-built requests
(pl_ortho_make_requests_range inside the timed region).

Reports tiles/s (CUDA events around the sweep on the launching stream), the kernel's own time per launch
(pl_timing: events around each launch) against the HBM roofline -- algorithmic bytes per tile =
196*196*4 written + 100*100*4 of the parent quadrant read = 193 664 -- a sha1 over all tiles, the same
levels 0..3 hashes the parity tests pin, and the oracle timed on the host cores for a bounded sample.

    python tools/ortho_sweep.py [--max-level 7] [--reps 3] [--cpu-level 4]
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "proland-4.0_b200"))
sys.path.insert(0, ROOT)
import proland_b200 as pl
import bench

W = 196
TILE_BYTES = W * W * 4 + (W // 2 + 2) ** 2 * 4


def run(ctx, torch, stream, sc, max_level, reps, peak, peak_kind):
    off = [(4 ** l - 1) // 3 for l in range(max_level + 2)]
    total = off[max_level + 1]
    pool = ctx.pool(pl.POOL_ORTHO, W, total)
    reqbuf = np.zeros(4 ** max_level, pl.ORTHO_REQ_DTYPE)

    def sweep():
        for l in range(max_level + 1):
            reqs = pl.ortho_make_requests_range(sc, l, 0, 4 ** l, out_slot0=off[l], parent_slot0=off[l - 1] if l else 0,
                                                out=reqbuf)
            ctx.ortho_batch(sc, pool, None, reqs)
    sweep()
    ctx.sync()
    ctx.timing_collect()
    ctx.timing_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(reps):
        sweep()
    e1.record(stream)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps
    ms = e0.elapsed_time(e1) / reps
    k_ms, launches, tiles = ctx.timing_collect()["ortho"]
    ctx.timing_enable(False)
    # the deepest level alone: one launch of 4^L tiles
    gbs = TILE_BYTES * tiles / (k_ms * 1e-3) / 1e9
    h = hashlib.sha1()
    first = []
    for s in range(total if total <= 5461 else 5461):
        t = pool.download(s)
        h.update(t.tobytes())
        if s < 85:
            first.append(hashlib.sha1(t.tobytes()).hexdigest()[:16])
    # the same sweep with the requests generated on the device (pl_ortho_produce_range)
    def sweep_dev():
        for l in range(max_level + 1):
            ctx.ortho_produce_range(sc, pool, l, 0, 4 ** l, off[l], off[l - 1] if l else 0, 0)
    sweep_dev()
    ctx.sync()
    e0.record(stream)
    for _ in range(reps):
        sweep_dev()
    e1.record(stream)
    torch.cuda.synchronize()
    ms_dev = e0.elapsed_time(e1) / reps
    pool.close()
    return {"tiles_per_sweep": total, "tiles_per_s": total / (ms * 1e-3), "ms_per_sweep": ms, "wall_ms_per_sweep": wall * 1e3,
            "device_requests": {"tiles_per_s": total / (ms_dev * 1e-3), "ms_per_sweep": ms_dev},
            "kernel": {"launches_per_sweep": launches // reps, "ms_per_sweep": k_ms / reps,
                       "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                                    "bytes_per_tile": TILE_BYTES, "peak_kind": peak_kind}},
            "sha1_levels_0_6": h.hexdigest()[:16], "levels_0_3_sha1": first}


def run_residuals(ctx, torch, stream, peak, peak_kind, level=6, reps=3):
    """ortho residual files on the device: 4^level TIFF/DEFLATE blobs of 196 x 196 x 3 bytes (64 distinct synthetic
    tiles, zlib level 6, cycled) through pl_ortho_decode_batch, then that level produced WITH its residuals
    (algorithmic bytes per tile: + 196*196*4 of residual read)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import resid_synth as rs
    rng = np.random.default_rng(20240612)
    distinct = []
    for _ in range(64):
        base = rng.integers(-6, 7, (W // 4 + 1, W // 4 + 1, 3))
        t = 128 + np.kron(base, np.ones((4, 4, 1), np.int64))[:W, :W] + rng.integers(-3, 4, (W, W, 3))
        distinct.append(t.astype(np.uint8))
    blobs64 = [rs.ortho_tiff_blob(t, 6) for t in distinct]
    n = 4 ** level
    blobs = [blobs64[i % 64] for i in range(n)]
    in_bytes = sum(len(b) for b in blobs)
    off = [(4 ** l - 1) // 3 for l in range(level + 2)]
    pool = ctx.pool(pl.POOL_ORTHO, W, off[level + 1])
    rpool = ctx.pool(pl.POOL_ORTHO, W, n)
    out = {}
    assert ctx.ortho_decode(rpool, blobs, list(range(n))) == 3
    ctx.timing_collect()
    ctx.timing_enable(True)
    for _ in range(reps):
        ctx.ortho_decode(rpool, blobs, list(range(n)))
    ms = ctx.timing_collect()["residual"][0] / reps
    ctx.timing_enable(False)
    assert np.array_equal(rpool.download(5)[..., :3], distinct[5])
    out["decode"] = {"tiles": n, "tiles_per_s": n / (ms * 1e-3), "ms_per_batch": ms,
                     "compressed_MBps": in_bytes / (ms * 1e-3) / 1e6, "decoded_GBps": n * W * W * 3 / (ms * 1e-3) / 1e9,
                     "compression_ratio": n * W * W * 3 / in_bytes}
    for name, hsv in (("hsv", 1), ("plain", 0)):
        sc = pl.ortho_scene(channels=3, hsv=hsv, cnoise=(70, 80, 100), rnoise=(60, 150, 20), noise_amp=[255] * 17, face=1)
        for l in range(level):
            ctx.ortho_batch(sc, pool, None, pl.ortho_make_requests_range(sc, l, 0, 4 ** l, out_slot0=off[l],
                                                                         parent_slot0=off[l - 1] if l else 0))
        reqs = pl.ortho_make_requests_range(sc, level, 0, n, out_slot0=off[level], parent_slot0=off[level - 1])
        reqs["resid_slot"] = np.arange(n)
        ctx.ortho_batch(sc, pool, rpool, reqs)
        ctx.sync()
        ctx.timing_collect()
        ctx.timing_enable(True)
        for _ in range(reps):
            ctx.ortho_batch(sc, pool, rpool, reqs)
        k_ms = ctx.timing_collect()["ortho"][0] / reps
        ctx.timing_enable(False)
        by = TILE_BYTES + W * W * 4
        gbs = by * n / (k_ms * 1e-3) / 1e9
        out[name] = {"tiles": n, "tiles_per_s": n / (k_ms * 1e-3), "ms_per_batch": k_ms,
                     "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                                  "bytes_per_tile": by, "peak_kind": peak_kind}}
    pool.close()
    rpool.close()
    return out


def main():
    import torch
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-level", type=int, default=7)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cpu-level", type=int, default=4)
    a = ap.parse_args()
    peak, peak_kind = bench.peaks()
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "ortho.json")))
    out = {"workload": "ortho tiles, face 1, levels 0..%d, 196-texel RGBA8 tiles" % a.max_level}
    with pl.Context(0) as ctx:
        stream = torch.cuda.Stream()
        ctx.set_stream(stream.cuda_stream)
        ctx.ortho_noise_init(W)
        hsv = pl.ortho_scene(hsv=1, cnoise=(70, 80, 100), rnoise=(60, 150, 20), noise_amp=[255] * 17, face=1)
        plain = pl.ortho_scene(hsv=0, cnoise=(127.5, 0, 0, 0), noise_amp=[0] + [255] * 16, face=3)
        out["terrain3_hsv"] = run(ctx, torch, stream, hsv, a.max_level, a.reps, peak, peak_kind)
        out["plain"] = run(ctx, torch, stream, plain, a.max_level, a.reps, peak, peak_kind)
        # the same scenes on a storage without alpha (terrain3's is RGB8): the alpha channel is not computed
        for name, kw in (("terrain3_hsv_rgb8", dict(hsv=1, cnoise=(70, 80, 100), rnoise=(60, 150, 20), noise_amp=[255] * 17, face=1)),
                         ("plain_rgb8", dict(hsv=0, cnoise=(127.5, 0, 0, 0), noise_amp=[0] + [255] * 16, face=3))):
            r = run(ctx, torch, stream, pl.ortho_scene(out_channels=3, **kw), a.max_level, a.reps, peak, peak_kind)
            r.pop("levels_0_3_sha1")
            out[name] = r
        out["with_residuals"] = run_residuals(ctx, torch, stream, peak, peak_kind)
    out["terrain3_hsv"]["golden_ok"] = out["terrain3_hsv"].pop("levels_0_3_sha1") == golden["terrain3_hsv"]["levels_0_3_sha1"]
    out["plain"]["golden_ok"] = out["plain"].pop("levels_0_3_sha1") == golden["plain"]["levels_0_3_sha1"]
    # CPU baseline: the oracle (a port, OpenMP over the tiles of a level) on a bounded sample
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    n = (4 ** (a.cpu_level + 1) - 1) // 3
    kw = dict(W=W, face=1, noise_amp=[255] * 17, noise_color=list(hsv.noise_color),
              root_noise_color=list(hsv.root_noise_color), hsv=1, scale=2.0)
    orc.ortho_quadtree(1, **kw)
    t0 = time.perf_counter()
    orc.ortho_quadtree(a.cpu_level, **kw)
    dt = time.perf_counter() - t0
    out["cpu_baseline"] = {"value": n / dt, "unit": "tiles/s", "cores": os.cpu_count(), "kind": "port",
                           "sample": "terrain3 hsv scene, levels 0..%d (%d tiles), oracle with OpenMP" % (a.cpu_level, n)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
