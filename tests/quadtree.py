"""Shared helpers for the parity tests: run a quadtree through the CUDA library
and through the oracle on the same parameters.  TEST INFRASTRUCTURE."""
import numpy as np


def level_tiles(level):
    """Row-major tiles of one level."""
    n = 1 << level
    return [(level, tx, ty) for ty in range(n) for tx in range(n)]


def gpu_quadtree(plb, ctx, max_level, *, noise_amp, face=0, root_quad_size=100000.0, flip=0,
                 noise_mode=1, no_clamp=0, sphere=0, elev_filter=1, want_stats=1, tiles_of=None):
    """Produce levels 0..max_level (all tiles) on the GPU; returns dict
    (level,tx,ty) -> (elev[101,101,3], norm[97,97,2], (zmin,zmax))."""
    W = 101
    tiles_of = tiles_of or level_tiles
    total = sum(len(tiles_of(l)) for l in range(max_level + 1))
    elev = ctx.pool(plb.POOL_ELEV, W, total)
    norm = ctx.pool(plb.POOL_NORM2, W - 4, total)
    ctx.noise_init(W)
    es = plb.elev_scene(W, 24, flip, noise_mode, no_clamp, want_stats)
    ns = plb.norm_scene(W - 4, 24, 2, elev_filter, 1, sphere)
    slot_of = {}
    for level in range(max_level + 1):
        tiles = tiles_of(level)
        reqs = plb.elev_make_reqs(tiles, tile_w=W, root_quad_size=root_quad_size,
                                  noise_amp=noise_amp, face=face)
        nreqs = plb.norm_make_reqs(tiles, ns, root_quad_size=root_quad_size)
        for i, t in enumerate(tiles):
            slot_of[t] = len(slot_of)
            reqs["out_slot"][i] = slot_of[t]
            if level > 0:
                reqs["parent_slot"][i] = slot_of[(level - 1, t[1] // 2, t[2] // 2)]
            nreqs["out_slot"][i] = slot_of[t]
            nreqs["elev_slot"][i] = slot_of[t]
        ctx.elevation_batch(es, elev, reqs)
        ctx.normal_batch(ns, norm, elev, nreqs)
    ctx.sync()
    keys = list(slot_of)
    stats = ctx.elev_stats(elev, [slot_of[k] for k in keys])
    return {k: (elev.download(slot_of[k]), norm.download(slot_of[k]), tuple(stats[i]))
            for i, k in enumerate(keys)}


def oracle_quadtree(orc, max_level, *, noise_amp, face=0, root_quad_size=100000.0, flip=0,
                    noise_mode=1, no_clamp=0, sphere=0, elev_filter=1, tiles_of=None):
    tiles_of = tiles_of or level_tiles
    scene = orc.make_scene(W=101, gridMeshSize=24, rootQuadSize=root_quad_size, face=face,
                           flip=flip, noise_mode=noise_mode, no_clamp=no_clamp,
                           noiseAmp=noise_amp, sphere=sphere, elev_filter=elev_filter)
    noise = orc.dem_noise(101)
    out = {}
    for level in range(max_level + 1):
        for t in tiles_of(level):
            parent = out[(level - 1, t[1] // 2, t[2] // 2)][0] if level > 0 else None
            e, n = orc.produce_pair(scene, noise, level, t[1], t[2], parent)
            out[t] = (e, n, orc.tile_minmax(e))
    return out


def compare(gpu, ref):
    """Returns (n_tiles, max |dh|, normal byte mismatches, stats mismatches)."""
    max_dh, nbad, sbad = 0.0, 0, 0
    for k, (e, n, s) in ref.items():
        ge, gn, gs = gpu[k]
        max_dh = max(max_dh, float(np.max(np.abs(ge.astype(np.float64) - e))))
        nbad += int(np.count_nonzero(gn != n))
        sbad += int(gs[0] != s[0]) + int(gs[1] != s[1])
    return len(ref), max_dh, nbad, sbad
