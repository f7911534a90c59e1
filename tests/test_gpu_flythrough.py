"""Config 5 (camera fly-through with LRU tile pools) on the GPU: the per-frame machinery of
tools/flythrough.py, with tiles sampled along the way checked bit for bit against the oracle."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu

AMP = [0] * 11 + [5, 2.5, 1, 0.5, 0.25, 0.1, 0.05, 0.025, 0.01, 0.01, 0.005, 0.005]


@pytest.mark.parametrize("asynchronous", [False, True])
def test_flythrough_tiles_match_the_oracle(oracle, tmp_path, asynchronous):
    import flythrough as ft
    import proland_host as ph
    if not os.path.exists(ph.LIB_PATH):
        ph.build()
    data = ft.write_residuals(str(tmp_path), face=2, max_level=5)
    res = oracle.Resid(data, delta=2)
    scene_o = oracle.make_scene(W=101, rootQuadSize=2 * ft.R, face=2, flip=1, noiseAmp=AMP, sphere=1, elev_filter=0, resid=res)
    noise = oracle.dem_noise(101)
    ref = {}

    def make(level, tx, ty):
        if (level, tx, ty) not in ref:
            parent = make(level - 1, tx // 2, ty // 2)[0] if level else None
            rt = res.create_tile(level, tx // 2, ty // 2) if res.has_tile(level, tx // 2, ty // 2) else None
            ref[(level, tx, ty)] = oracle.produce_pair(scene_o, noise, level, tx, ty, parent, rt)
        return ref[(level, tx, ty)]

    rng = np.random.default_rng(3)
    checked = []

    def on_frame(k, cam, terrain, elevations, normals):
        if k % 6 != 5:
            return
        leaves = [q for q in terrain.quads() if q[3]]
        for idx in rng.choice(len(leaves), size=2, replace=False):
            level, tx, ty, _ = leaves[idx]
            nt = normals.find_tile(level, tx, ty, include_cache=True, done=True)
            et = elevations.find_tile(level, tx, ty, include_cache=True, done=True)
            if nt is None or et is None:
                assert asynchronous          # async: a leaf may still be on its way
                continue
            e, n = make(level, tx, ty)
            assert np.array_equal(et.download(), e), (k, level, tx, ty)
            assert np.array_equal(nt.download(), n), (k, level, tx, ty)
            checked.append((level, tx, ty))

    out = ft.run(frames=36, asynchronous=asynchronous, ntiles=1296, max_level=9, data_dir=str(tmp_path), on_frame=on_frame)
    assert len(checked) >= (3 if asynchronous else 8) and max(c[0] for c in checked) >= 5
    assert out["tiles_made"] > 500 and out["quads_max"] < 1296
    # batching: far fewer kernel launches than tiles
    assert out["tiles_per_launch"] > 3.0
    assert 0.0 < out["miss_rate"]["groundNormals"] <= 1.0


def test_tile_sampler_z_feeds_the_quadtree(oracle, tmp_path):
    """TileSamplerZ (TileSamplerZ.cpp:253-386): the z range of every tile the sampler holds reaches TerrainQuad zmin /
    zmax, the zm texel under the camera reaches TerrainNode::groundHeightAtCamera two read-backs late, and the quadtree's
    split rule consumes it: with the camera low over mountains the tree is the one a model fed with the live ground
    heights predicts, not the one of ground height 0."""
    import math
    import flythrough as ft
    import proland_host as ph
    import proland_b200 as plb
    if not os.path.exists(ph.LIB_PATH):
        ph.build()
    amp = "-3250,-400,-200,-100,-50,-25,10,5"      # level 0 dominates: a mountain of the root tile stays one below
    xml = ft.archive(2500, False, 2).replace(ft.NOISE, amp).replace('residuals="groundResiduals2" ', "").replace('flip="true"', 'flip="false"')
    ft.write_residuals(str(tmp_path), face=2, max_level=3)
    scene = ph.Scene(xml, data_dir=str(tmp_path))
    elevations = scene.producer("groundElevations2")
    sched = scene.scheduler("defaultScheduler")
    SIZE, MAXL = 20000.0, 9      # a 40 km root quad: a few hundred metres of ground height decide splits at levels 8, 9
    terrain = ph.Terrain(SIZE, zmin=0.0, zmax=10000.0, split_factor=2.0, max_level=MAXL)
    sz = ph.SamplerZ("elevationSampler", elevations)
    split = ph.lib().plh_split_distance(2.0, 1024.0, math.radians(80.0))
    ph.ground_height(reset=True)
    scene_o = oracle.make_scene(W=101, rootQuadSize=2 * SIZE, face=2, flip=0, noiseAmp=[float(a) for a in amp.split(",")],
                                sphere=1, elev_filter=0)
    noise = oracle.dem_noise(101)
    ref = {}

    def make(level, tx, ty):
        if (level, tx, ty) not in ref:
            parent = make(level - 1, tx // 2, ty // 2) if level else None
            ref[(level, tx, ty)] = oracle.produce_pair(scene_o, noise, level, tx, ty, parent, None, want_normals=False)[0]
        return ref[(level, tx, ty)]

    def model_tree(cam, ground, max_level=MAXL):
        """TerrainQuad::update with a given ground height (tests/test_host_terrain.py holds the rule itself)"""
        out = []

        def rec(level, tx, ty, ox, oy, l):
            zlo, zhi = min(0.0, ground), max(0.0, ground)
            dist = max(abs(cam[2] - zhi), min(abs(cam[0] - ox), abs(cam[0] - ox - l)), min(abs(cam[1] - oy), abs(cam[1] - oy - l)))
            split = np.float32(dist) < np.float32(l) * np.float32(split_dist) and level < max_level
            out.append((level, tx, ty, int(not split)))
            if split:
                hl = float(np.float32(l) / np.float32(2.0))
                for c in range(4):
                    rec(level + 1, 2 * tx + (c & 1), 2 * ty + (c >> 1), ox + (c & 1) * hl, oy + (c >> 1) * hl, hl)
        rec(0, 0, 0, -SIZE, -SIZE, 2.0 * SIZE)
        return out

    split_dist = split
    # find a high spot of the level-0 tile and fly low over it
    e0 = make(0, 0, 0)
    j, i = np.unravel_index(np.argmax(e0[2:-2, 2:-2, 2]), (97, 97))
    peak = float(e0[2 + j, 2 + i, 2])
    assert peak > 1000.0
    cx, cy = -SIZE + (i + 0.5) / 96.0 * 2 * SIZE, -SIZE + (j + 0.5) / 96.0 * 2 * SIZE
    grounds = []
    for k in range(60):     # 16 tiles are read back per frame: enough frames to drain the needReadback set
        cam = (cx + 0.3 * k, cy + 0.2 * k, 750.0)        # moves > 0.1 m per frame: the camera texel is read each frame
        g_before = ph.ground_height()[0]
        terrain.update(*cam, split_dist=split)
        assert terrain.quads() == model_tree(cam, g_before), k
        ph.frame_update(sched, terrain, [sz])
        grounds.append(ph.ground_height())
    issued, applied, waiting = sz.counts()
    assert issued >= 30 and applied >= issued - 2 and waiting == 0, (issued, applied, waiting)
    # the ground height: the zm texel under the camera of the deepest finished quad, i.e. a height of the oracle's tiles
    g, gn = grounds[-1]
    assert g > 100.0 and gn > 100.0, (g, gn, peak, grounds[:12])
    # two read-backs late: 0 until the first camera read-back is applied, then `next` leads `current` by one
    assert grounds[0] == (0.0, 0.0) and any(a == 0.0 and b > 0.0 for a, b in grounds[:8])
    held = {}
    for level, tx, ty, zmin, zmax in terrain.quads_z():
        held[(level, tx, ty)] = (zmin, zmax)
    # every quad whose tile the sampler holds has the oracle's z range over [2, W-3]^2 (TileSamplerZ.cpp:60-64)
    checked = 0
    for (level, tx, ty), (zmin, zmax) in held.items():
        t = elevations.find_tile(level, tx, ty, include_cache=False, done=True)
        if t is None or (zmin, zmax) == (0.0, 10000.0):
            continue
        zm = make(level, tx, ty)[2:-2, 2:-2, 2]
        assert zmin == zm.min() and zmax == zm.max(), (level, tx, ty)
        checked += 1
    assert checked >= 20
    # the camera height above ground: with ground height 0 the tree would be shallower
    assert len(model_tree(cam, 0.0)) < len(terrain.quads())
    assert max(q[0] for q in held) > max(q[0] for q in model_tree(cam, 0.0)), "the live ground height deepens the tree under the camera"
    # and the value itself is the oracle's: the camera texel of some level of the chain under the camera
    txs = [(l, int((cam[0] + SIZE) / (2 * SIZE) * (1 << l)), int((cam[1] + SIZE) / (2 * SIZE) * (1 << l))) for l in range(MAXL + 1)]
    cand = set()
    for l, tx, ty in txs:
        e = make(l, tx, ty)
        for kk in range(max(0, k - 4), k + 1):
            c = (cx + 0.3 * kk, cy + 0.2 * kk)
            fx = (c[0] + SIZE) / (2 * SIZE) * (1 << l) - tx
            fy = (c[1] + SIZE) / (2 * SIZE) * (1 << l) - ty
            if 0 <= fx < 1 and 0 <= fy < 1:
                cand.add(float(e[min(int(np.float32(fy) * 96), 95) + 2, min(int(np.float32(fx) * 96), 95) + 2, 2]))
    assert g in cand and gn in cand
    sz.close()
    terrain.close()
    scene.close()
