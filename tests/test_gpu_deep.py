"""GPU parity at the DEPTH the bench runs: full-tile, bit-for-bit compares of elevation + normal tiles against the
oracle at the deepest levels of every BASELINE config, through pl_pair_batch (the fused kernel the bench times).

  config 2  demo-fractalplanet : 11 random level-10 leaves per cube face (66 tiles on each of levels 8, 9, 10 --
            the slope / curvature noise levels -- plus all their ancestors), sphere normals, LINEAR storage
  config 3  demo-earth-srtm    : 64 random level-12 leaves on one face, int16 residual tiles on ~70 % of the chain
            tiles, flipped diagonals, NEAREST storage, sphere normals
  config 4  subtree sweep      : 64 random level-14 leaves of the flat fractal terrain

Every tile of every chain is compared (not only the leaves): 2 000+ full tiles."""
import numpy as np
import pytest

import glsl_cases as gc

pytestmark = pytest.mark.gpu
W = 101


def _chain_tiles(leaves):
    by_level = {}
    for leaf in leaves:
        for (l, tx, ty) in gc.chain(*leaf):
            by_level.setdefault(l, set()).add((l, tx, ty))
    return [sorted(by_level[l]) for l in sorted(by_level)]


def _deep(plb, ctx, oracle, leaves, *, amp, face, rqs, sphere, flip=0, filt=1, resid_prob=0.0, seed=0, arith=0):
    levels = _chain_tiles(leaves)
    total = sum(len(t) for t in levels)
    rng = np.random.default_rng(seed)
    elev = ctx.pool(plb.POOL_ELEV, W, total)
    norm = ctx.pool(plb.POOL_NORM2, W - 4, total)
    ctx.noise_init(W)
    es = plb.elev_scene(W, 24, flip, 1, 0, 1)
    ns = plb.norm_scene(W - 4, 24, 2, filt, 1, sphere)
    if arith:
        ns.arith = arith
    noise = oracle.dem_noise(W)
    strict_bad = [0]

    def oracle_pair(l, tx, ty, parent, rt):
        p = oracle.elev_uniforms(l, tx, ty, rootQuadSize=rqs, noiseAmp=amp, face=face, flip=flip, noise_mode=1,
                                 has_resid=int(rt is not None), resid_W=197 if rt is not None else 0)
        e = oracle.upsample_tile(p, parent, rt, noise)
        q = oracle.normal_uniforms(l, tx, ty, rootQuadSize=rqs, sphere=sphere, elev_filter=filt)
        n = oracle.pack_unorm8(oracle.normal_tile(q, e), 2)
        if arith:   # how far the reference's own text, read without contraction, is from the canonical reading
            ns_ = oracle.pack_unorm8(oracle.normal_tile(q, e, L=oracle.strict()), 2)
            strict_bad[0] += int(np.count_nonzero(ns_ != n))
        return e, n
    # residual tiles (197 wide, one per 2 x 2 elevation tiles: earth-srtm's mod = 2), int16 pool
    rtiles, rslot = {}, {}
    if resid_prob > 0:
        for tl in levels:
            for (l, tx, ty) in tl:
                k = (l, tx // 2, ty // 2)
                if l >= 1 and k not in rtiles and rng.random() < resid_prob:
                    a = max(1.0, 300.0 / (1 << l))
                    rtiles[k] = np.round(rng.normal(0, a, (197, 197))).astype(np.int16)
        rpool = ctx.pool(plb.POOL_RESID_I16, 197, max(1, len(rtiles)))
        for s, (k, t) in enumerate(sorted(rtiles.items())):
            rslot[k] = s
            rpool.upload(s, t)
    else:
        rpool = None
    slot_of, ref = {}, {}
    for tl in levels:
        has = [(l, tx // 2, ty // 2) in rtiles for (l, tx, ty) in tl]
        reqs = plb.elev_make_reqs(tl, tile_w=W, root_quad_size=rqs, noise_amp=amp, face=face,
                                  resid_tile_w=197 if rpool else 0, has_resid=has if rpool else None)
        nreqs = plb.norm_make_reqs(tl, ns, root_quad_size=rqs)
        for i, t in enumerate(tl):
            l, tx, ty = t
            slot_of[t] = len(slot_of)
            reqs["out_slot"][i] = nreqs["out_slot"][i] = nreqs["elev_slot"][i] = slot_of[t]
            if l:
                reqs["parent_slot"][i] = slot_of[(l - 1, tx // 2, ty // 2)]
            if has[i]:
                reqs["resid_slot"][i] = rslot[(l, tx // 2, ty // 2)]
            parent = ref[(l - 1, tx // 2, ty // 2)][0] if l else None
            rt = rtiles[(l, tx // 2, ty // 2)].astype(np.float32) if has[i] else None
            ref[t] = oracle_pair(l, tx, ty, parent, rt)
        ctx.pair_batch(es, ns, elev, norm, reqs, nreqs, resid=rpool)
    ctx.sync()
    stats = ctx.elev_stats(elev, [slot_of[t] for t in ref])
    bad_bytes = worst = nbytes = 0
    for i, (t, (e, n)) in enumerate(ref.items()):
        ge = elev.download(slot_of[t])
        assert np.array_equal(ge, e), ("elevation tile differs from the oracle", face, t)
        assert tuple(stats[i]) == oracle.tile_minmax(e), t
        gn = norm.download(slot_of[t])
        if arith == 0:
            assert np.array_equal(gn, n), ("normal tile differs from the oracle", face, t)
        else:
            d = np.abs(gn.astype(int) - n.astype(int))
            worst = max(worst, int(d.max()))
            bad_bytes += int(np.count_nonzero(d))
            nbytes += d.size
    return len(ref), worst, bad_bytes, nbytes, strict_bad[0]


def _leaves(rng, level, n):
    return [(level, int(rng.integers(0, 1 << level)), int(rng.integers(0, 1 << level))) for _ in range(n)]


@pytest.mark.parametrize("face", [1, 2, 3, 4, 5, 6])
def test_fractalplanet_levels_8_9_10(plb, ctx, oracle, face):
    rng = np.random.default_rng(100 + face)
    leaves = _leaves(rng, 10, 11)
    # one leaf on each edge of the face: the noise-layer selection's cube-edge branches (ElevationProducer.cpp:348-366)
    leaves[0] = (10, 0, leaves[0][2])
    leaves[1] = (10, 1023, leaves[1][2])
    leaves[2] = (10, leaves[2][1], 0)
    leaves[3] = (10, leaves[3][1], 1023)
    n, *_ = _deep(plb, ctx, oracle, leaves, amp=gc.PLANET, face=face, rqs=12720000.0, sphere=1)
    assert n >= 11 * 3 + 20


def test_earth_srtm_level_12(plb, ctx, oracle):
    rng = np.random.default_rng(12)
    n, *_ = _deep(plb, ctx, oracle, _leaves(rng, 12, 64), amp=gc.SRTM, face=2, rqs=12720000.0, sphere=1, flip=1, filt=0,
                  resid_prob=0.7, seed=3)
    assert n >= 64 * 6


def test_subtree_sweep_level_14(plb, ctx, oracle):
    rng = np.random.default_rng(14)
    n, *_ = _deep(plb, ctx, oracle, _leaves(rng, 14, 64), amp=gc.FRACTAL + [0, 0, 0], face=0, rqs=100000.0, sphere=0)
    assert n >= 64 * 8
