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
