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


def test_fractalterrain_levels_0_4(plb, ctx, oracle):
    """config 1 (demo-fractalterrain): variant D, clamp, flat RG8 normals, LINEAR storage."""
    assert _run(plb, ctx, oracle, 4, noise_amp=FRACTAL) == 341


@pytest.mark.parametrize("noise_mode,flip,no_clamp", [(0, 0, 0), (1, 0, 1), (0, 1, 0), (1, 1, 0)])
def test_shader_variants(plb, ctx, oracle, noise_mode, flip, no_clamp):
    """variants A (plain), B/D-noClamp, C (plain+flip), D+flip."""
    _run(plb, ctx, oracle, 3, noise_amp=FRACTAL[:2] + [30, 20], noise_mode=noise_mode, flip=flip,
         no_clamp=no_clamp)


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
