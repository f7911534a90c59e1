"""Driver-run records of BASELINE.json's other configs (bench.py calls these at N = 1 after the headline measurement and
puts them under `configs`; each is also runnable on its own: python tools/configs.py [1 3 4 5]).

  config 1  demo-fractalterrain: flat face, levels 0..8, both arithmetic contracts (tools/fractalterrain.py)
  config 3  demo-earth-srtm shape: residual decode (device-resident archive and host blobs), the fused kernel's
            residual variants on the int16 and the float pool, and decode + production back to back
  config 4  full-subtree sweeps at level 14, d = 6, 8, 10 (tools/subtree_sweep.py)
  config 5  camera fly-through through the C++ host layer (tools/flythrough.py): tiles/s, p99 frame time

Every record carries its own `roofline` (algorithmic bytes of the dominant kernel / its CUDA-event time; `traffic` = the
ncu-measured DRAM bytes per pair of the matching capture in profiles/, or null).  Timings are CUDA events on the context's
stream; nothing here runs under a profiler.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "proland-4.0_b200"), os.path.join(ROOT, "tests"), ROOT, os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)

SRTM_AMP = [0] * 7 + [5, 2.5, 1]          # earth-srtm.xml: no noise down to the residual levels, then a little
PLANET_SIZE = 12720000.0
RESID_PAIR_BYTES = {"int16": 212500, "float": 232902}      # 192 098 + the residual window as stored


def _events(torch, stream):
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def config1(pl, ctx, torch, stream):
    import fractalterrain as ft
    out = {}
    for arith, name in ((pl.ARITH_FAST, "fast"), (pl.ARITH_EXACT, "exact")):
        out[name] = ft.run(pl, ctx, torch, stream, max_level=8, reps=5, arith=arith)
    return out


def config3(pl, ctx, torch, stream, peak, peak_kind, n_decode=32768, level=8):
    import resid_synth as rs
    rng = np.random.default_rng(20240612)
    distinct = [rs.fractal_tile(rng, 197, 40) for _ in range(64)]
    blobs64 = [rs.tiff_blob(t, 6) for t in distinct]
    sizes64 = np.array([len(b) for b in blobs64], np.uint32)
    offs64 = np.concatenate([[0], np.cumsum((sizes64[:-1] + 15) & ~np.uint32(15), dtype=np.uint64)]).astype(np.uint64)
    arch = bytearray(int(offs64[-1] + sizes64[-1]))
    for o, b in zip(offs64, blobs64):
        arch[int(o):int(o) + len(b)] = b
    out = {"workload": "config 3 demo-earth-srtm shape: 197 x 197 int16 residual tiles (64 distinct synthetic SRTM-shaped "
                       "tiles, zlib level 6, cycled), sphere face 2, flip, NEAREST elevation storage",
           "compression_ratio": float(64 * 197 * 197 * 2 / sizes64.sum())}
    ctx.noise_init(101)
    # ---- decode: the archive resident in HBM (the reference maps the file once), tiles located by offset
    store = ctx.blobs(bytes(arch))
    idx = np.arange(n_decode) % 64
    pool = ctx.pool(pl.POOL_RESID_I16, 197, n_decode)
    args = (pool, store, offs64[idx], sizes64[idx], [197] * n_decode, np.arange(n_decode, dtype=np.int32))
    ctx.residual_decode_stored(*args)
    ctx.sync()
    ctx.timing_collect()
    ctx.timing_enable(True)
    e0, e1 = _events(torch, stream)
    t0 = time.perf_counter()
    e0.record(stream)
    reps = 3
    for _ in range(reps):
        ctx.residual_decode_stored(*args)
    e1.record(stream)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps
    k_ms = ctx.timing_collect()["residual"][0] / reps
    ctx.timing_enable(False)
    assert np.array_equal(pool.download(n_decode - 1)[:197, :197], distinct[(n_decode - 1) % 64])
    out["decode_resident"] = {
        "tiles": n_decode, "tiles_per_s_kernels": n_decode / (k_ms * 1e-3), "tiles_per_s_call": n_decode / wall,
        "kernel_ms": k_ms, "decoded_GBps": n_decode * 197 * 197 * 2 / (k_ms * 1e-3) / 1e9,
        "compressed_GBps": float(sizes64[idx].sum()) / (k_ms * 1e-3) / 1e9,
        "path": "pl_residual_decode_stored: tokenizer (lane per stream) + resolver (warp per stream) kernels, int16 pool",
        "roofline": {"bound": "latency", "note": "a DEFLATE stream is one serial chain: the tokenizer is bound by the "
                     "per-symbol dependency chain times the streams resident per SM, not by HBM (DESIGN 3.3)",
                     "achieved": n_decode * (197 * 197 * 2 + float(sizes64.mean())) / (k_ms * 1e-3) / 1e9, "peak": peak,
                     "unit": "GB/s", "frac": n_decode * (197 * 197 * 2 + float(sizes64.mean())) / (k_ms * 1e-3) / 1e9 / peak,
                     "peak_kind": peak_kind, "traffic": None}}
    pool.close()
    # ---- decode from HOST blobs (packing + upload inside the call), the warp-per-stream decoder's batch size
    nh = 4096
    blobs = [blobs64[i % 64] for i in range(nh)]
    hpool = ctx.pool(pl.POOL_RESID_I16, 197, nh)
    ctx.residual_decode(hpool, blobs, [197] * nh, list(range(nh)))
    ctx.sync()
    t0 = time.perf_counter()
    ctx.residual_decode(hpool, blobs, [197] * nh, list(range(nh)))
    ctx.sync()
    out["decode_host_blobs"] = {"tiles": nh, "tiles_per_s_call": nh / (time.perf_counter() - t0),
                                "path": "pl_residual_decode_batch: blobs packed and uploaded inside the call"}
    # ---- the fused kernel with residuals, one level-`level` batch; then decode + production back to back
    off = [sum(4 ** k for k in range(l)) for l in range(level + 2)]
    n = 4 ** level
    nres = n // 4
    elev = ctx.pool(pl.POOL_ELEV, 101, off[level + 1])
    norm = ctx.pool(pl.POOL_NORM2, 97, off[level + 1])
    for arith, aname in ((pl.ARITH_FAST, "fast"), (pl.ARITH_EXACT, "exact")):
        sc = pl.sweep_scene(noise_amp=SRTM_AMP, face=2, root_quad_size=PLANET_SIZE, sphere=1, flip=1,
                            elev_filter=pl.FILTER_NEAREST, want_stats=1, arith=arith)
        sc.elev.resid_scale = 1.0
        for l in range(level):      # ancestors: fractal only
            ctx.produce_range(sc, elev, norm, l, 0, 4 ** l, off[l], off[l - 1] if l else 0, 0)
        ids = pl.make_tile_ids_range(level, 0, n, off[level], off[level - 1], 0)
        # one residual tile per 2 x 2 tiles
        ids["resid_slot"] = (ids["tx"] // 2) + (ids["ty"] // 2) * (1 << (level - 1))
        ridx = np.arange(nres) % 64
        for kind, name in ((pl.POOL_RESID_I16, "int16"), (pl.POOL_RESID_F32, "float")):
            rpool = ctx.pool(kind, 197, nres)
            dargs = (rpool, store, offs64[ridx], sizes64[ridx], [197] * nres, np.arange(nres, dtype=np.int32))
            ctx.residual_decode_stored(*dargs)
            ctx.pair_batch_ids(sc, elev, norm, ids, resid=rpool)      # warm-up
            ctx.sync()
            ctx.timing_collect()
            ctx.timing_enable(True)
            reps = 5
            for _ in range(reps):
                ctx.pair_batch_ids(sc, elev, norm, ids, resid=rpool)
            ms = ctx.timing_collect()["pair"][0] / reps
            gbs = RESID_PAIR_BYTES[name] * n / (ms * 1e-3) / 1e9
            rec = {"pairs": n, "pairs_per_s": n / (ms * 1e-3), "kernel_ms": ms,
                   "roofline": {"bound": "hbm", "kernel": "pair (residual variant)", "achieved": gbs, "peak": peak, "unit": "GB/s",
                                "frac": gbs / peak, "bytes_per_pair": RESID_PAIR_BYTES[name], "peak_kind": peak_kind, "traffic": None}}
            # decode INSIDE the timed region: the residual tiles of the batch, then the batch
            e0, e1 = _events(torch, stream)
            ctx.timing_collect()
            e0.record(stream)
            for _ in range(reps):
                ctx.residual_decode_stored(*dargs)
                ctx.pair_batch_ids(sc, elev, norm, ids, resid=rpool)
            e1.record(stream)
            torch.cuda.synchronize()
            kt = ctx.timing_collect()
            ctx.timing_enable(False)
            tot = e0.elapsed_time(e1) / reps
            rec["with_decode"] = {"pairs_per_s": n / (tot * 1e-3), "ms": tot, "residual_tiles": nres,
                                  "decode_kernel_ms": kt["residual"][0] / reps, "pair_kernel_ms": kt["pair"][0] / reps,
                                  "fraction_of_kernel_only": ms / tot}
            out["pairs_%s_pool_%s" % (name, aname)] = rec
            rpool.close()
    norm.close()
    elev.close()
    hpool.close()
    try:
        out["planet_sweep"] = config3_sweep(pl, ctx, torch, stream, peak, peak_kind, store, offs64, sizes64)
    except Exception as ex:
        out["planet_sweep"] = {"error": "%s: %s" % (type(ex).__name__, ex)}
    try:
        out["every_tile_with_residuals_level9"] = config3_every_tile_large(pl, ctx, torch, stream, peak, peak_kind, store, offs64, sizes64)
    except Exception as ex:
        out["every_tile_with_residuals_level9"] = {"error": "%s: %s" % (type(ex).__name__, ex)}
    store.close()
    return out


def config3_every_tile_large(pl, ctx, torch, stream, peak, peak_kind, store, offs64, sizes64, level=9, reps=2):
    """The worst case at a batch size the two-kernel decoder is built for: EVERY tile of one level-`level` batch of a sphere face
    reads a residual window (one 197 x 197 int16 tile per 2 x 2 elevation tiles), its residual tiles decoded inside the timed
    region, int16 pool, PL_ARITH_FAST."""
    off = [sum(4 ** k for k in range(l)) for l in range(level + 2)]
    n = 4 ** level
    nres = n // 4
    elev = ctx.pool(pl.POOL_ELEV, 101, off[level + 1])
    norm = ctx.pool(pl.POOL_NORM2, 97, off[level + 1])
    rpool = ctx.pool(pl.POOL_RESID_I16, 197, nres)
    sc = pl.sweep_scene(noise_amp=SRTM_AMP, face=2, root_quad_size=PLANET_SIZE, sphere=1, flip=1,
                        elev_filter=pl.FILTER_NEAREST, want_stats=1, arith=pl.ARITH_FAST)
    sc.elev.resid_scale = 1.0
    for l in range(level):      # ancestors: fractal only
        ctx.produce_range(sc, elev, norm, l, 0, 4 ** l, off[l], off[l - 1] if l else 0, 0)
    ids = pl.make_tile_ids_range(level, 0, n, off[level], off[level - 1], 0)
    ids["resid_slot"] = (ids["tx"] // 2) + (ids["ty"] // 2) * (1 << (level - 1))
    ridx = np.arange(nres) % len(sizes64)
    dargs = (rpool, store, offs64[ridx], sizes64[ridx], [197] * nres, np.arange(nres, dtype=np.int32))
    ctx.residual_decode_stored(*dargs)
    ctx.pair_batch_ids(sc, elev, norm, ids, resid=rpool)
    ctx.sync()
    e0, e1 = _events(torch, stream)
    ctx.timing_collect()
    ctx.timing_enable(True)
    e0.record(stream)
    for _ in range(reps):
        ctx.residual_decode_stored(*dargs)
        ctx.pair_batch_ids(sc, elev, norm, ids, resid=rpool)
    e1.record(stream)
    torch.cuda.synchronize()
    kt = ctx.timing_collect()
    ctx.timing_enable(False)
    tot = e0.elapsed_time(e1) / reps
    pair_ms, dec_ms = kt["pair"][0] / reps, kt["residual"][0] / reps
    gbs = RESID_PAIR_BYTES["int16"] * n / (pair_ms * 1e-3) / 1e9
    rec = {"workload": "every one of the %d level-%d tiles of a sphere face reads a residual window; its %d residual tiles (197 x 197 "
                       "int16, zlib 6) decoded inside the timed region by the two-kernel decoder" % (n, level, nres),
           "pairs": n, "pairs_per_s": n / (tot * 1e-3), "ms": tot, "residual_tiles": nres, "decode_kernel_ms": dec_ms,
           "pair_kernel_ms": pair_ms, "decode_tiles_per_s": nres / (dec_ms * 1e-3), "fraction_of_kernel_only": pair_ms / tot,
           "roofline": {"bound": "hbm", "kernel": "pair (residual variant)", "achieved": gbs, "peak": peak, "unit": "GB/s",
                        "frac": gbs / peak, "bytes_per_pair": RESID_PAIR_BYTES["int16"], "peak_kind": peak_kind, "traffic": None}}
    rpool.close()
    norm.close()
    elev.close()
    return rec


def config3_sweep(pl, ctx, torch, stream, peak, peak_kind, store, offs64, sizes64, max_level=10, resid_levels=6, reps=2):
    """Config 3 as SURVEY 8d defines it: the cube-face quadtree of a planet whose residual files end at a finite level
    (stored level 8 = elevation levels 0..6), fractal amplification below.  The residual tiles of the whole sweep (one per
    2 x 2 elevation tiles, 1366 per face) are decoded in ONE batch INSIDE the timed region, then the 6 faces are swept
    breadth-first through the recycling pool (plan: sweep.py); tiles of levels 0..resid_levels read their residual window."""
    import sweep as plan
    _, cap = plan.region_offsets(max_level)
    elev = ctx.pool(pl.POOL_ELEV, 101, cap)
    norm = ctx.pool(pl.POOL_NORM2, 97, cap)
    level_base = [0, 1]
    for l in range(2, resid_levels + 2):
        level_base.append(level_base[-1] + 4 ** (l - 2))
    per_face = level_base[resid_levels + 1]
    nres = 6 * per_face
    rpool = ctx.pool(pl.POOL_RESID_I16, 197, nres)
    ridx = np.arange(nres) % len(sizes64)
    dargs = (rpool, store, offs64[ridx], sizes64[ridx], [197] * nres, np.arange(nres, dtype=np.int32))
    amp = SRTM_AMP + [0.5] * max(0, max_level + 1 - len(SRTM_AMP))
    scenes = {}
    for f in range(1, 7):
        sc = pl.sweep_scene(noise_amp=amp, face=f, root_quad_size=PLANET_SIZE, sphere=1, flip=1, elev_filter=pl.FILTER_NEAREST,
                            want_stats=1, arith=pl.ARITH_FAST)
        sc.elev.resid_scale = 1.0
        scenes[f] = sc
    units = plan.planet_units()
    pairs = plan.pairs_in_units(units, max_level, True)
    id_buf = np.zeros(4 ** max(resid_levels, 1), pl.TILE_ID_DTYPE)
    resid_pairs = 0

    def sweep():
        nonlocal resid_pairs
        resid_pairs = 0
        ctx.residual_decode_stored(*dargs)
        for f, level, m0, n, s0, p0, pm0 in plan.batches(units, max_level):
            if level <= resid_levels:
                ids = pl.make_tile_ids_range(level, m0, n, s0, p0, pm0, out=id_buf)
                ids["resid_slot"] = (f - 1) * per_face + level_base[level] + ((m0 + np.arange(n)) >> 2 if level else 0)
                ctx.pair_batch_ids(scenes[f], elev, norm, ids, resid=rpool)
                resid_pairs += n
            else:
                ctx.produce_range(scenes[f], elev, norm, level, m0, n, s0, p0, pm0)

    sweep()
    ctx.sync()
    ctx.timing_collect()
    ctx.timing_enable(True)
    e0, e1 = _events(torch, stream)
    e0.record(stream)
    for _ in range(reps):
        sweep()
    e1.record(stream)
    torch.cuda.synchronize()
    kt = ctx.timing_collect()
    ctx.timing_enable(False)
    ms = e0.elapsed_time(e1) / reps
    pair_ms = kt["pair"][0] / reps
    dec_ms = kt["residual"][0] / reps
    nbytes = (pairs - resid_pairs) * 192098 + resid_pairs * RESID_PAIR_BYTES["int16"]
    st = ctx.elev_stats_readback_end(ctx.elev_stats_readback_begin(elev, 0, plan.ROOT_SLOTS))
    rec = {"workload": "config 3 as a planet sweep: 6 cube faces, levels 0..%d, %d pairs; residual files end at elevation level %d "
                       "(%d pairs read a residual window, %d residual tiles of 197 x 197 int16 decoded in one batch inside the timed "
                       "region), fractal amplification below; flip, NEAREST elevation storage, sphere normals, PL_ARITH_FAST"
                       % (max_level, pairs, resid_levels, resid_pairs, nres),
           "pairs": pairs, "pairs_per_s": pairs / (ms * 1e-3), "ms_per_sweep": ms, "decode_kernel_ms": dec_ms,
           "pair_kernel_ms": pair_ms, "decode_tiles_per_s": nres / (dec_ms * 1e-3) if dec_ms else None,
           "fraction_of_kernel_only": pair_ms / ms,
           "roofline": {"bound": "hbm", "kernel": "pair", "achieved": nbytes / (pair_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": nbytes / (pair_ms * 1e-3) / 1e9 / peak, "frac_with_decode_in_the_denominator": nbytes / (ms * 1e-3) / 1e9 / peak,
                        "bytes_per_pair": nbytes / pairs, "peak_kind": peak_kind, "traffic": None},
           "fingerprint": {"root_stats_sum": float(st.astype(np.float64).sum())}}
    rpool.close()
    norm.close()
    elev.close()
    return rec


def config4(pl, ctx, torch, stream, peak, peak_kind, ds=(6, 8, 10)):
    """pools are created before the timed region (a sweep recycles one pool for all its units): what is timed is the
    production of the sweep, its per-unit statistics read-backs included"""
    import subtree_sweep as ss
    import sweep
    import bench
    out = {}
    ctx.noise_init(101)
    for d in ds:
        plan = sweep.SubtreeSweep(14 - d, (1 << (14 - d)) // 3, (1 << (14 - d)) // 5, d, 7)
        pools = (ctx.pool(pl.POOL_ELEV, 101, plan.capacity), ctx.pool(pl.POOL_NORM2, 97, plan.capacity))
        rec = {}
        for arith, aname in ((pl.ARITH_FAST, "fast"), (pl.ARITH_EXACT, "exact")):
            sub = {}
            for levels, key in ((False, "launch_per_level"), (True, "launch_per_subtree")):
                kw = dict(scene_kw=dict(arith=arith))
                ss.run_sweep(pl, ctx, d, pools=pools, levels=levels, **kw)      # warm-up
                ctx.sync()
                e0, e1 = _events(torch, stream)
                ctx.timing_collect()
                ctx.timing_enable(True)
                l0 = ctx.launches
                reps = 5 if d <= 8 else 2
                e0.record(stream)
                for _ in range(reps):
                    n, fp, _ = ss.run_sweep(pl, ctx, d, pools=pools, levels=levels, **kw)
                e1.record(stream)
                torch.cuda.synchronize()
                kt = ctx.timing_collect()
                ctx.timing_enable(False)
                ms = e0.elapsed_time(e1) / reps
                pair_ms, launches, tiles = kt["pair"]
                gbs = bench.PAIR_BYTES * tiles / (pair_ms * 1e-3) / 1e9
                sub[key] = {"pairs_per_s": n / (ms * 1e-3), "ms": ms, "launches": int((ctx.launches - l0) // reps),
                            "roofline": {"bound": "hbm", "kernel": "pair", "achieved": gbs, "peak": peak, "unit": "GB/s",
                                         "frac": gbs / peak, "bytes_per_pair": bench.PAIR_BYTES, "peak_kind": peak_kind,
                                         "share_of_sweep": pair_ms / reps / ms, "traffic_per_pair": None},
                            "fingerprint_xor": fp["xor"]}
            sub["identical"] = bool(sub["launch_per_level"]["fingerprint_xor"] == sub["launch_per_subtree"]["fingerprint_xor"])
            rec[aname] = sub
        # the elevation statistics are bit-exact under both contracts: one fingerprint
        rec["identical"] = bool(rec["fast"]["identical"] and rec["exact"]["identical"] and
                                rec["fast"]["launch_per_level"]["fingerprint_xor"] == rec["exact"]["launch_per_level"]["fingerprint_xor"])
        rec["workload"] = ("all level-14 descendants of one level-%d tile (+ ancestors): %d pairs, flat face 0, both arithmetic contracts; "
                           "launch_per_level = one pl_produce_range per level, launch_per_subtree = pl_produce_levels (the chain / a "
                           "unit in one launch, parent -> child dependency resolved inside the kernel)" % (14 - d, n))
        rec["pairs"] = n
        out["d%d" % d] = rec
        pools[0].close()
        pools[1].close()
    return out


def config5(frames=200):
    import flythrough as fl
    rec = fl.run(frames=frames, asynchronous=True)
    sync = fl.run(frames=frames, asynchronous=False)
    rec["synchronous"] = {k: sync[k] for k in ("tiles_made", "tiles_per_s", "frame_ms_p50", "frame_ms_p99", "tiles_per_launch")}
    return rec


def run_all(pl, ctx, torch, stream, peak, peak_kind, which=(1, 3, 4, 5)):
    out = {}
    for k, fn in ((1, lambda: config1(pl, ctx, torch, stream)), (3, lambda: config3(pl, ctx, torch, stream, peak, peak_kind)),
                  (4, lambda: config4(pl, ctx, torch, stream, peak, peak_kind)), (5, config5)):
        if k not in which:
            continue
        t0 = time.perf_counter()
        try:
            out["config%d" % k] = fn()
        except Exception as ex:     # a sub-record must not take the headline line with it
            out["config%d" % k] = {"error": "%s: %s" % (type(ex).__name__, ex)}
        ctx.sync()
        out["config%d" % k]["seconds"] = time.perf_counter() - t0
    return out


if __name__ == "__main__":
    import torch
    import proland_b200 as pl
    import bench
    which = tuple(int(a) for a in sys.argv[1:]) or (1, 3, 4, 5)
    peak, kind = bench.peaks()
    with pl.Context(0) as ctx:
        stream = torch.cuda.Stream()
        ctx.set_stream(stream.cuda_stream)
        with torch.cuda.stream(stream):
            print(json.dumps(run_all(pl, ctx, torch, stream, peak, kind, which), indent=1))
