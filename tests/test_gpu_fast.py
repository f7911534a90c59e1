"""PL_ARITH_FAST: the tolerance contract of the normal pass (include/proland_b200.h, pl_normal_tile.cuh).

Elevation tiles stay bit-identical to the oracle; a normal byte may differ from the canonical evaluation by at most
ONE unorm8 step (127.5 * dn = 1 -> at most 0.45 degrees of the encoded normal).  How MANY bytes differ is bounded
against the reference itself: the reference's shader text read without contraction (the strict oracle ==
oracle/_ref/libref_glsl.so bit for bit, tests/test_glsl_pin.py) is an equally admissible evaluation of the same GLSL,
and FAST must be no farther from the canonical reading than 1.5 x that reading is (+ 1e-4 of the bytes), and below
4e-3 of the bytes in any case.  Why not a flat 1e-3: over flat water (zm = max(zf, 0) = 0) the tangent-frame normal
is (0, 0, 1) up to the patch curvature, so r = g = 127.5 +- 0.1 -- every texel near the tile's centre lines sits ON the
unorm8 rounding tie, and ANY two fp32 evaluations (strict vs canonical included) disagree on ~50 of an ocean tile's
18 818 bytes (2.8e-3).  Checked on the deep chains of every BASELINE config (the same cases test_gpu_deep.py holds bit-exact under
PL_ARITH_EXACT), through the fused kernel, the separate normal kernel and the device-generated sweep."""
import numpy as np
import pytest

import glsl_cases as gc
from test_gpu_deep import _deep, _leaves

pytestmark = pytest.mark.gpu
FAST = 1


def _check(n, worst, bad, nbytes, strict_bad):
    assert worst <= 1, "a normal byte is %d unorm8 steps from the oracle" % worst
    assert bad <= 1.5 * strict_bad + 1e-4 * nbytes, \
        "%d of %d normal bytes differ; the reference's non-contracted reading differs on %d" % (bad, nbytes, strict_bad)
    assert bad < 4e-3 * nbytes, "%d of %d normal bytes differ (%.2e)" % (bad, nbytes, bad / nbytes)


@pytest.mark.parametrize("face", [1, 3, 6])
def test_fast_fractalplanet_deep(plb, ctx, oracle, face):
    rng = np.random.default_rng(200 + face)
    leaves = _leaves(rng, 10, 11)
    leaves[0] = (10, 0, leaves[0][2])
    leaves[1] = (10, 1023, 1023)
    _check(*_deep(plb, ctx, oracle, leaves, amp=gc.PLANET, face=face, rqs=12720000.0, sphere=1, arith=FAST))


def test_fast_earth_srtm_level_12(plb, ctx, oracle):
    rng = np.random.default_rng(12)
    _check(*_deep(plb, ctx, oracle, _leaves(rng, 12, 32), amp=gc.SRTM, face=2, rqs=12720000.0, sphere=1, flip=1, filt=0,
                  resid_prob=0.7, seed=3, arith=FAST))


def test_fast_flat_terrains(plb, ctx, oracle):
    rng = np.random.default_rng(14)
    _check(*_deep(plb, ctx, oracle, _leaves(rng, 14, 24), amp=gc.FRACTAL + [0, 0, 0], face=0, rqs=100000.0, sphere=0,
                  arith=FAST))
    _check(*_deep(plb, ctx, oracle, _leaves(rng, 8, 24), amp=gc.FRACTAL, face=0, rqs=100000.0, sphere=0, filt=0,
                  arith=FAST))


@pytest.mark.parametrize("fused", [True, False])
def test_fast_sweep_levels_0_5(plb, ctx, oracle, fused):
    """pl_produce_range with arith = FAST (device-generated requests; fused kernel or the two passes): levels 0..5 of
    a planet face -- the levels below R/64 keep the exact position code, only the normalisation is approximate"""
    import quadtree as qt
    kw = dict(noise_amp=gc.PLANET, face=5, root_quad_size=12720000.0, sphere=1)
    sc = plb.sweep_scene(want_stats=1, arith=plb.ARITH_FAST, **kw)
    max_level = 5
    off = [sum(4 ** k for k in range(l)) for l in range(max_level + 2)]
    elev = ctx.pool(plb.POOL_ELEV, 101, off[-1])
    norm = ctx.pool(plb.POOL_NORM2, 97, off[-1])
    ctx.noise_init(101)
    ctx.no_fuse(not fused)
    for l in range(max_level + 1):
        ctx.produce_range(sc, elev, norm, l, 0, 4 ** l, off[l], off[l - 1] if l else 0, 0)
    ctx.sync()
    ctx.no_fuse(False)
    ref = qt.oracle_quadtree(oracle, max_level, **kw)
    worst = bad = nbytes = sbad = 0
    for (l, tx, ty), (e, n, s) in ref.items():
        slot = off[l] + plb.morton_encode(tx, ty)
        assert np.array_equal(elev.download(slot), e), (l, tx, ty)
        d = np.abs(norm.download(slot).astype(int) - n.astype(int))
        worst, bad, nbytes = max(worst, int(d.max())), bad + int(np.count_nonzero(d)), nbytes + d.size
        q = oracle.normal_uniforms(l, tx, ty, rootQuadSize=12720000.0, sphere=1, elev_filter=1)
        sbad += int(np.count_nonzero(oracle.pack_unorm8(oracle.normal_tile(q, e, L=oracle.strict()), 2) != n))
    _check(len(ref), worst, bad, nbytes, sbad)


@pytest.mark.parametrize("sphere,level", [(0, 8), (1, 8), (1, 7), (1, 6)])
def test_slim_layout_is_byte_identical_to_the_regular_one(plb, ctx, sphere, level):
    """launches whose tiles all take the register form run the fused kernel in its slim layout (no second window copy, no
    position planes: 4 CTAs per SM; by default on flat scenes, forced here on spheres too): the same elevation planes,
    statistics and normal bytes as the regular layout, on a whole level with slope noise, through produce_range, the
    identity entry point and host-built requests.  Planet level 7 is the first whose quads are R/64 wide (smoothstep factor
    exactly 1: the host's per-launch criterion says "register form"), level 6 the last that is not (the slim layout must
    not be chosen: both runs take the regular one)"""
    n = 1024
    kw = dict(noise_amp=gc.PLANET, face=3, root_quad_size=12720000.0, sphere=1) if sphere else \
        dict(noise_amp=gc.FRACTAL + [2, 1, 1], face=0, root_quad_size=100000.0, sphere=0)
    sc = plb.sweep_scene(want_stats=1, arith=plb.ARITH_FAST, **kw)
    ctx.noise_init(101)
    elev = ctx.pool(plb.POOL_ELEV, 101, 2 * n + n // 4)
    norm = ctx.pool(plb.POOL_NORM2, 97, 2 * n + n // 4)
    ctx.no_slim(1)
    ctx.produce_range(sc, elev, norm, level - 1, 0, n // 4, 2 * n, 0, 0)          # parents (from zero grandparents)
    got = {}
    for mode, base in ((1, 0), (-1, n)):                                           # regular layout, then slim
        ctx.no_slim(mode)
        ctx.produce_range(sc, elev, norm, level, 0, n, base, 2 * n, 0)
        ctx.sync()
        got[mode] = ([elev.download(base + s) for s in range(0, n, 37)], [norm.download(base + s) for s in range(0, n, 37)],
                     ctx.elev_stats_range(elev, base, n))
    for a, b in zip(got[1][0], got[-1][0]):
        assert np.array_equal(a, b)
    for a, b in zip(got[1][1], got[-1][1]):
        assert np.array_equal(a, b)
    assert np.array_equal(got[1][2], got[-1][2])
    # the identity entry point takes the same decision per batch
    ids = plb.make_tile_ids_range(level, 0, n, n, 2 * n, 0)
    ids["norm_slot"] = np.arange(n)                                                # normals over the regular run's slots
    ctx.pair_batch_ids(sc, elev, norm, ids)
    ctx.sync()
    for k, s in enumerate(range(0, n, 37)):
        assert np.array_equal(norm.download(s), got[1][1][k])
        assert np.array_equal(elev.download(n + s), got[1][0][k])
    ctx.no_slim(0)
