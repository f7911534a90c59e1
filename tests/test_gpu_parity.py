"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the
same seeds.  Elevations and normals are compared for EXACT equality: kernels
and oracle share one canonical fp32 evaluation order (oracle/orc_fp.h), so any
difference at all is a bug.  The distance to the other admissible GLSL reading
(no contraction) is measured in test_oracle.py and is the stated tolerance."""
import numpy as np
import pytest

import quadtree as qt

pytestmark = pytest.mark.gpu

FRACTAL = [-140, -100, -15, -8, 5, 2.5, 1.5, 1, 0.5, 0.25, 0.1, 0.05]
PLANET = [-3250, -1590, -1125, -795, -561, -397, -140, -100, 15, 8, 5, 2.5, 1.5, 1, 0.5, 0.25, 0.1, 0.05]


def _run(plb, ctx, oracle, max_level, **kw):
    gpu = qt.gpu_quadtree(plb, ctx, max_level, **kw)
    ref = qt.oracle_quadtree(oracle, max_level, **kw)
    n, max_dh, nbad, sbad = qt.compare(gpu, ref)
    assert max_dh == 0.0, "elevations differ from the oracle: max |dh| = %g" % max_dh
    assert nbad == 0, "%d normal bytes differ from the oracle" % nbad
    assert sbad == 0, "%d zmin/zmax values differ" % sbad
    return n


def test_fpexact_matches_ieee(plb, ctx):
    """the branch-free division / reciprocal / sqrt of pl_fpexact.cuh are the IEEE results,
    bit for bit, on the domain the tile path lives in (and at zero for sqrt / numerators)."""
    rng = np.random.default_rng(7)
    n = 1 << 22
    mant = rng.random(n, np.float32) + np.float32(1.0)
    a = (mant * np.exp2(rng.integers(-60, 60, n)).astype(np.float32)
         * rng.choice(np.array([-1, 1], np.float32), n))
    b = ((rng.random(n, np.float32) + np.float32(1.0)) * np.exp2(rng.integers(-30, 30, n)).astype(np.float32)
         * rng.choice(np.array([-1, 1], np.float32), n))
    a[:1000] = 0.0                                    # exact zeros: flat terrain
    a[1000:2000] = np.float32(2.0) ** -100            # lower domain bound of sqrt
    b[2000:3000] = np.float32(100000.0 / 96.0) / np.exp2(np.arange(1000) % 20).astype(np.float32)
    out = ctx.fpexact(a, b)
    for k, name in ((0, "div"), (2, "rcp"), (4, "sqrt")):
        bad = np.flatnonzero(out[k].view(np.uint32) != out[k + 1].view(np.uint32))
        # +0 / -0 compare equal as values
        bad = bad[out[k][bad] != out[k + 1][bad]]
        assert bad.size == 0, "%s: %d mismatches, first a=%r b=%r" % (name, bad.size, a[bad[0]], b[bad[0]])


def test_fractalterrain_levels_0_4(plb, ctx, oracle):
    """config 1 (demo-fractalterrain): variant D, clamp, flat RG8 normals, LINEAR storage."""
    assert _run(plb, ctx, oracle, 4, noise_amp=FRACTAL) == 341


@pytest.mark.parametrize("noise_mode,flip,no_clamp", [(0, 0, 0), (1, 0, 1), (0, 1, 0), (1, 1, 0)])
def test_shader_variants(plb, ctx, oracle, noise_mode, flip, no_clamp):
    """variants A (plain), B/D-noClamp, C (plain+flip), D+flip."""
    _run(plb, ctx, oracle, 3, noise_amp=FRACTAL[:2] + [30, 20], noise_mode=noise_mode, flip=flip,
         no_clamp=no_clamp)


def test_generic_geometry_kernels(plb, ctx, oracle):
    """the runtime-geometry kernels (used for tile sizes other than 101/97) on the same case"""
    ctx.force_generic(True)
    _run(plb, ctx, oracle, 3, noise_amp=FRACTAL[:2] + [30, 20], flip=1)


@pytest.mark.parametrize("face", [1, 2, 5, 6])
def test_fractalplanet_faces(plb, ctx, oracle, face):
    """config 2 (demo-fractalplanet): cube faces, sphere-deformed normals."""
    _run(plb, ctx, oracle, 3, noise_amp=PLANET, face=face, root_quad_size=12720000.0, sphere=1)


def test_sphere_deep_chain(plb, ctx, oracle):
    """one root-to-level-9 chain on a sphere face: exercises smoothstep == 1 (level >= 7),
    positive slope-modulated noise (level >= 8) and NEAREST elevation storage."""
    def chain(level):
        return [(level, 397 >> (9 - level), 341 >> (9 - level))]
    _run(plb, ctx, oracle, 9, noise_amp=PLANET, face=3, root_quad_size=12720000.0, sphere=1,
         elev_filter=0, tiles_of=chain)


@pytest.mark.parametrize("face,level", [(0, 5), (1, 6), (4, 6), (6, 7)])
def test_device_requests_match_host(plb, ctx, face, level):
    """pl_produce_range generates the per-tile uniforms on the device; they must be
    byte-identical to the host's (integer decisions through cnoise, fp64 geometry)."""
    sc = plb.sweep_scene(noise_amp=PLANET, face=face, root_quad_size=12720000.0, sphere=1)
    n = min(4 ** level, 4096)
    m0 = (4 ** level - n) // 4 * 4 if level > 5 else 0
    elev = ctx.pool(plb.POOL_ELEV, 101, n + n // 4 + 8)
    norm = ctx.pool(plb.POOL_NORM2, 97, n + n // 4 + 8)
    ctx.noise_init(101)
    p0 = n   # parents live after the outputs
    ctx.produce_range(sc, elev, norm, level, m0, n, 0, p0, m0 >> 2)
    de, dn = ctx.last_requests(n)
    he, hn = plb.make_requests_range(sc, level, m0, n, 0, p0, m0 >> 2)
    assert de.tobytes() == he.tobytes()
    assert dn.tobytes() == hn.tobytes()


def test_produce_range_equals_per_tile_path(plb, ctx, oracle):
    """the device-driven Morton sweep and the per-tile request path give the same tiles."""
    amp = PLANET
    kw = dict(noise_amp=amp, face=2, root_quad_size=12720000.0, sphere=1)
    sc = plb.sweep_scene(want_stats=1, **kw)
    max_level = 4
    total = sum(4 ** l for l in range(max_level + 1))
    elev = ctx.pool(plb.POOL_ELEV, 101, total)
    norm = ctx.pool(plb.POOL_NORM2, 97, total)
    ctx.noise_init(101)
    off = [sum(4 ** k for k in range(l)) for l in range(max_level + 1)]
    for l in range(max_level + 1):
        ctx.produce_range(sc, elev, norm, l, 0, 4 ** l, off[l], off[l - 1] if l else 0, 0)
    ctx.sync()
    ref = qt.oracle_quadtree(oracle, max_level, **kw)
    for (l, tx, ty), (e, n, s) in ref.items():
        slot = off[l] + plb.morton_encode(tx, ty)
        assert np.array_equal(elev.download(slot), e), (l, tx, ty)
        assert np.array_equal(norm.download(slot), n), (l, tx, ty)
