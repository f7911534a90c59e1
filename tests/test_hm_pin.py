"""The residual builder's oracle pinned on the reference's own builder.

oracle/_ref/libref_hm.so is preprocess/terrain/{Preprocess, HeightMipmap, AbstractTileCache, ColorMipmap, ApertureMipmap,
Util}.cpp and util/mfs.cpp of the reference checkout compiled UNCHANGED over shims for Ork's headers and for libtiff
(oracle/ref_shim/hm; libtiff is a binary-only dependency, its files are kept in memory).  Running
proland::preprocessSphericalDem / preprocessDem from it on the maps of tests/hm_cases.py gives residual files; the
restatement oracle/orc_preprocess.c (cube projections, SphericalHeightFunction / PlaneHeightFunction, setCube
stitching, getTileHeight's corner and edge rules, decimated mipmap levels, computeResidual / encodeResidual /
computeApproxTile, the level-0 and constant-tile rules) must reproduce every int16 sample of every tile of every file.
Integer work: the bar is bit-exact."""
import json
import os

import pytest

import hm_cases as hc

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "hm.json")))["cases"]


@pytest.mark.parametrize("name", sorted(hc.CASES))
def test_oracle_builder_equals_committed_reference_digests(oracle, name):
    """runs everywhere (no reference checkout needed)"""
    assert hc.digest(hc.oracle_record(oracle, name)) == GOLDEN[name]


@pytest.mark.parametrize("name", sorted(hc.CASES))
def test_reference_builder_equals_committed_digests_and_the_oracle_tile_for_tile(oracle, name):
    """the golden file is what oracle/_ref/libref_hm.so produces today, and the oracle agrees tile for tile"""
    if oracle.hm() is None:
        pytest.skip("oracle/_ref/libref_hm.so not built (reference checkout absent)")
    ref = hc.reference_record(oracle, name)
    assert hc.digest(ref) == GOLDEN[name]
    assert hc.first_difference(ref, hc.oracle_record(oracle, name)) is None


def test_cube_edges_agree_between_neighbouring_faces(oracle):
    """setCube's purpose: the border samples a face reads across an edge are the neighbour's interior samples, so the
    two faces' height tiles agree on the shared strip (the oracle's stitching, every level, all 12 edges)"""
    name = "sphere_12_48_l2_scale2"
    faces = hc.base_grids(oracle, name)
    B, _, max_level = hc.levels(name)
    for level in range(max_level + 1):
        n = 1 + (B >> (max_level - level))
        for f in range(6):
            for k in range(3, n - 3):
                # one step outside each edge equals some face's interior sample one step inside its own edge
                for (x, y) in ((-1, k), (n, k), (k, -1), (k, n)):
                    h = oracle.hm_height(faces, max_level, level, f, x, y)
                    cands = {oracle.hm_height(faces, max_level, level, g, a, b)
                             for g in range(6) if g != f for (a, b) in ((1, k), (n - 2, k), (k, 1), (k, n - 2),
                                                                        (1, n - 1 - k), (n - 2, n - 1 - k), (n - 1 - k, 1), (n - 1 - k, n - 2))}
                    assert h in cands
