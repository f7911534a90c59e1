"""The float part of the oracle pinned on the reference's own shader text.

oracle/_ref/libref_glsl.so is upsampleShader.glsl (demo + terrain1/2/4 variants), normalShader.glsl (demo +
terrain1/2) and upsampleOrthoShader.glsl (demo + examples) of the reference checkout, compiled UNCHANGED as C++
behind oracle/ref_shim/glsl_shim.h (plain IEEE fp32, one rounding per GLSL operation).  The restatement
oracle/orc_{elevation,normal,ortho}.c read without contraction (liborc_strict.so) must reproduce it BIT FOR BIT;
the canonical reading (liborc.so: a*b+c fused, the order the CUDA kernels share) then differs from the reference
text only by the contraction GLSL 3.30 leaves to the implementation -- measured here and stated as the tolerance:
    elevations  max |dh| <= 1e-5 x (height range of the tile set)
    normals     <= 1 unorm8 step (0.5 degree), on < 2 % of the bytes of a tile
    ortho       <= 1 unorm8 step on < 1e-4 of the bytes
"""
import json
import os

import numpy as np
import pytest

import glsl_cases as gc

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "glsl.json")))["cases"]


def test_strict_oracle_equals_committed_glsl_hashes(oracle):
    """runs everywhere (no reference checkout needed): every case of tests/glsl_cases.py, produced by the
    non-contracted restatement, hashes to what the reference's GLSL text produced (tests/golden/glsl.json)"""
    got = gc.run(oracle, gc.Engine.STRICT)
    assert set(got) == set(GOLDEN)
    bad = sorted(k for k in got if got[k] != GOLDEN[k])
    assert not bad, "%d of %d cases differ from the reference's GLSL, first: %s" % (len(bad), len(got), bad[:5])
    assert len(got) >= 400


def test_glsl_reference_build_equals_committed_hashes(oracle):
    """the golden file is what oracle/_ref/libref_glsl.so produces today (guards the shim and the fixture)"""
    if oracle.glsl() is None:
        pytest.skip("oracle/_ref/libref_glsl.so not built (reference checkout absent)")
    got = gc.run(oracle, gc.Engine.GLSL, full=False)
    bad = sorted(k for k in got if got[k] != GOLDEN[k])
    assert not bad, bad[:5]


def test_canonical_oracle_within_tolerance_of_the_glsl(oracle):
    """the canonical (fused) reading against the reference's text on the deep chain of every BASELINE config:
    the stated tolerances above, texel by texel"""
    if oracle.glsl() is None:
        pytest.skip("oracle/_ref/libref_glsl.so not built (reference checkout absent)")
    noise = oracle.dem_noise(101)
    runs = [("config 1", gc.FRACTAL, 0, 100000.0, 0, (8, 201, 77), gc.LINEAR),
            ("config 2", gc.PLANET, 3, 12720000.0, 0, (10, 750, 413), gc.LINEAR),
            ("config 3", gc.SRTM, 2, 12720000.0, 1, (12, 2901, 1717), gc.NEAREST),
            ("config 4", gc.FRACTAL + [0, 0, 0], 0, 100000.0, 0, (14, 9999, 12345), gc.LINEAR)]
    rng = np.random.default_rng(20240612)
    resid = np.round(rng.normal(0, 40, (197, 197))).astype(np.float32)
    for name, amp, face, rqs, flip, leaf, filt in runs:
        pc = pg = None
        lo, hi, max_dh, worst_frac, max_step = np.inf, -np.inf, 0.0, 0.0, 0
        for (l, tx, ty) in gc.chain(*leaf):
            has_resid = int(name == "config 3" and l >= 1)
            p = oracle.elev_uniforms(l, tx, ty, rootQuadSize=rqs, noiseAmp=amp, face=face, flip=flip, noise_mode=1,
                                     has_resid=has_resid, resid_W=197 if has_resid else 0)
            r = resid if has_resid else None
            ec = oracle.upsample_tile(p, pc, r, noise)
            eg = oracle.glsl_upsample_tile("D", p, pg, r, noise, filt)
            lo, hi = min(lo, float(eg[..., 0].min())), max(hi, float(eg[..., 0].max()))
            max_dh = max(max_dh, float(np.abs(ec - eg).max()))
            q = oracle.normal_uniforms(l, tx, ty, rootQuadSize=rqs, sphere=int(face != 0), elev_filter=filt)
            nc = oracle.pack_unorm8(oracle.normal_tile(q, ec), 2).astype(int)
            ng = oracle.pack_unorm8(oracle.glsl_normal_tile("demo", q, eg), 2).astype(int)
            max_step = max(max_step, int(np.abs(nc - ng).max()))
            worst_frac = max(worst_frac, float(np.count_nonzero(nc != ng)) / nc.size)
            pc, pg = ec, eg
        assert hi - lo > 50, name
        assert max_dh <= 1e-5 * (hi - lo), (name, max_dh, hi - lo)
        assert max_step <= 1 and worst_frac < 0.02, (name, max_step, worst_frac)


def test_linear_filter_with_exact_float_weights_stays_within_tolerance(oracle):
    """a LINEAR elevation storage read through a sampler that keeps full fp32 weights (a software rasteriser)
    instead of 8 subtexel bits (GPU texture units): the fp32 rounding of the texture coordinates leaks ~1e-6 of
    the neighbouring texel into every fetch; still inside the stated 1e-5 x range"""
    if oracle.glsl() is None:
        pytest.skip("oracle/_ref/libref_glsl.so not built (reference checkout absent)")
    noise = oracle.dem_noise(101)
    parent = None
    lo, hi, worst = np.inf, -np.inf, 0.0
    for (l, tx, ty) in gc.chain(6, 46, 25):
        p = oracle.elev_uniforms(l, tx, ty, rootQuadSize=12720000.0, noiseAmp=gc.PLANET, face=3, noise_mode=1)
        e8 = oracle.glsl_upsample_tile("D", p, parent, None, noise, gc.LINEAR, 8)
        e0 = oracle.glsl_upsample_tile("D", p, parent, None, noise, gc.LINEAR, 0)
        lo, hi = min(lo, float(e8[..., 0].min())), max(hi, float(e8[..., 0].max()))
        worst = max(worst, float(np.abs(e8 - e0).max()))
        parent = e8
    assert 0 < worst <= 1e-5 * (hi - lo), (worst, hi - lo)


def test_ortho_canonical_within_one_step_of_the_glsl(oracle):
    if oracle.glsl() is None:
        pytest.skip("oracle/_ref/libref_glsl.so not built (reference checkout absent)")
    W = 196
    noise = oracle.ortho_noise(W)
    kw = dict(W=W, face=1, noise_amp=[255] * 17, noise_color=[np.float32(v) / np.float32(255) for v in (70, 80, 100, 255)],
              root_noise_color=[np.float32(v) / np.float32(255) for v in (60, 150, 20, 127.5)], hsv=1, scale=2.0)
    parent, nbad, ntot = None, 0, 0
    for (l, tx, ty) in gc.chain(4, 11, 6):
        p = oracle.ortho_uniforms(l, tx, ty, **kw)
        tc = oracle.ortho_tile(p, parent, None, noise)
        tg = oracle.pack_unorm8(oracle.glsl_ortho_tile(0, p, parent, None, noise), 4)
        d = np.abs(tc.astype(int) - tg.astype(int))
        assert d.max() <= 1
        nbad += int(np.count_nonzero(d)); ntot += d.size
        parent = tc
    assert nbad < 1e-4 * ntot


def test_glsl_quadtree_driver_agrees_with_the_port(oracle):
    """bench.py's `reference_glsl` datum: a whole quadtree by the reference's shader text (ref_glsl_produce_quadtree, OpenMP)
    against the oracle port's -- same tile count, per-tile (zmin + zmax) checksum within the contraction tolerance"""
    if oracle.glsl() is None:
        pytest.skip("oracle/_ref/libref_glsl.so not built (reference checkout absent)")
    scene = oracle.make_scene(W=101, gridMeshSize=24, rootQuadSize=12720000.0, face=3, flip=0, noise_mode=1, no_clamp=0,
                              noiseAmp=gc.PLANET, sphere=1, elev_filter=1)
    n, cs, lo, hi = oracle.produce_quadtree(scene, 3, 2)
    got = oracle.glsl_produce_quadtree(scene, 3, 2)
    assert got is not None and got[0] == n == 85
    assert abs(got[1] - cs) <= 1e-6 * abs(cs) and abs(got[2] - lo) <= 1e-5 * (hi - lo) and abs(got[3] - hi) <= 1e-5 * (hi - lo)
