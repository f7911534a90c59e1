"""The residual builder on the device (HeightMipmap + preprocessDem / preprocessSphericalDem) against
the oracle and the digests of the files the reference's own builder writes (tests/golden/hm.json, tests/test_hm_pin.py).
Integer work: bit-exact -- every base sample, every height tile across the cube's edges and corners, every int16
residual of every file."""
import json
import os
import tempfile

import numpy as np
import pytest

import hm_cases as hc

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "hm.json")))["cases"]


@pytest.fixture(scope="module")
def ph():
    import proland_host
    if not os.path.exists(proland_host.LIB_PATH):
        proland_host.build()
    return proland_host


@pytest.mark.parametrize("name", sorted(hc.CASES))
def test_base_level_grids_equal_the_oracle(plb, ctx, oracle, name):
    """SphericalHeightFunction / PlaneHeightFunction + (short) h on the device: every sample of every face"""
    spherical = hc.CASES[name][0]
    B = hc.levels(name)[0]
    src = hc.source_map(name)
    cube = ctx.height_cube_from_latlon(src, B) if spherical else ctx.height_cube_from_plane(src, B)
    try:
        want = hc.base_grids(oracle, name)
        assert cube.nfaces == len(want)
        for f, w in enumerate(want):
            got = cube.download(f)
            assert np.array_equal(got, w), "face %d: %d of %d samples differ" % (f, int((got != w).sum()), w.size)
    finally:
        cube.close()


def test_undecidable_samples_are_redone_on_the_host(plb, ctx, oracle):
    """a map of exact integers on a lattice the cube samples hit exactly (the poles, the equator crossings) makes the
    truncation (short)(float) h sit on its decision boundary: those samples go through the host's libm and still
    equal the oracle"""
    sw, sh, B = 64, 32, 96
    yy, xx = np.mgrid[0:sh, 0:sw]
    src = (xx * 7 + yy * 13).astype(np.float32)
    before = plb.lib().pl_debug_height_unsure(ctx.h)
    cube = ctx.height_cube_from_latlon(src, B)
    try:
        redone = plb.lib().pl_debug_height_unsure(ctx.h) - before
        for f in range(6):
            assert np.array_equal(cube.download(f), oracle.spherical_base(src, f, B)), f
        assert redone > 0
    finally:
        cube.close()


@pytest.mark.parametrize("name", ["sphere_12_48_l2_scale2", "plane_24_96_l2"])
def test_height_tiles_across_edges_and_corners_equal_the_oracle(plb, ctx, oracle, name):
    """HeightMipmap::getTile with setCube's stitching (pl_height_tiles): every tile of every level of every face,
    from host-supplied base grids (pl_height_cube_create)"""
    _, _, _, minT, T, _, scale, _ = hc.CASES[name]
    B, _, max_level = hc.levels(name)
    faces = hc.base_grids(oracle, name)
    cube = ctx.height_cube(faces)
    try:
        tl = hc.tile_list(name)
        pool = ctx.pool(plb.POOL_RESID_F32, T + 5, len(tl) * len(faces))
        reqs = np.zeros(len(tl) * len(faces), plb.HEIGHT_REQ_DTYPE)
        k = 0
        for f in range(len(faces)):
            for (l, tx, ty, ts) in tl:
                reqs[k] = (f, l, tx, ty, k, (0, 0, 0))
                k += 1
        cube.tiles(pool, minT, T, reqs, scale)
        k = 0
        for f in range(len(faces)):
            for (l, tx, ty, ts) in tl:
                want = oracle.hm_get_tile(faces, max_level, minT, T, l, f, tx, ty, scale)
                got = pool.download(k)
                assert np.array_equal(got[:ts + 5, :ts + 5], want[:ts + 5, :ts + 5]), (f, l, tx, ty)
                k += 1
    finally:
        cube.close()


def test_height_tile_requests_are_validated(plb, ctx, oracle):
    faces = [np.zeros((97, 97), np.int16)]
    cube = ctx.height_cube(faces)
    try:
        pool = ctx.pool(plb.POOL_RESID_F32, 101, 2)
        for bad in [(1, 0, 0, 0, 0), (0, 3, 0, 0, 0), (0, 0, 1, 0, 0), (0, 0, 0, 0, 2), (0, -1, 0, 0, 0)]:
            reqs = np.zeros(1, plb.HEIGHT_REQ_DTYPE)
            reqs[0] = bad + ((0, 0, 0),)
            with pytest.raises(plb.PlError):
                cube.tiles(pool, 24, 96, reqs)
        with pytest.raises(plb.PlError):
            cube.tiles(pool, 24, 192, np.zeros(1, plb.HEIGHT_REQ_DTYPE))       # 197 does not fit a 101 pool
    finally:
        cube.close()


@pytest.mark.parametrize("name", sorted(hc.CASES))
def test_preprocess_dem_writes_the_files_the_reference_writes(ph, plb, oracle, name):
    """proland::preprocessDem / preprocessSphericalDem of the host layer, whole: source map -> DEM*.dat.  Header, blob
    sharing and every int16 of every tile equal the reference builder's files (committed digests), and -- when the
    reference build is here -- tile for tile with the position of the first difference"""
    spherical, _, _, minT, T, maxL, scale, _ = hc.CASES[name]
    with tempfile.TemporaryDirectory() as tmp:
        dst = os.path.join(tmp, "dst")
        ph.preprocess_dem(hc.source_map(name), minT, T, maxL, dst, scale, spherical)
        got = hc.record_of_files(oracle, name, dst)
        mtime = {f: os.path.getmtime(os.path.join(dst, f)) for f in os.listdir(dst)}
        assert sorted(mtime) == (["DEM%d.dat" % i for i in range(1, 7)] if spherical else ["DEM.dat"])
        # a second call finds the files and leaves them alone (Preprocess.cpp:516-518, 539-544)
        ph.preprocess_dem(hc.source_map(name), minT, T, maxL, dst, scale, spherical)
        assert mtime == {f: os.path.getmtime(os.path.join(dst, f)) for f in os.listdir(dst)}
    diff = hc.first_difference(got, hc.oracle_record(oracle, name))
    assert diff is None, diff
    assert hc.digest(got) == GOLDEN[name]
    if oracle.hm() is not None:
        assert hc.first_difference(got, hc.reference_record(oracle, name)) is None


def test_built_files_feed_the_residual_producer(ph, plb, ctx, oracle):
    """the loop closed through the run-time path: a file built here, read back by pl_residual_decode_batch, gives the
    builder's int16 tiles"""
    name = "sphere_24_96_l1"
    spherical, _, _, minT, T, maxL, scale, _ = hc.CASES[name]
    with tempfile.TemporaryDirectory() as tmp:
        dst = os.path.join(tmp, "dst")
        ph.preprocess_dem(hc.source_map(name), minT, T, maxL, dst, scale, spherical)
        rd = oracle.Resid(open(os.path.join(dst, "DEM3.dat"), "rb").read())
    rec = hc.oracle_record(oracle, name)[2]
    tl = hc.tile_list(name)
    own = [tid for tid in range(len(tl)) if rec["shared"][tid] == tid]
    pool = ctx.pool(plb.POOL_RESID_I16, T + 5, len(own))
    ctx.residual_decode(pool, [rd.blob(t) for t in own], [tl[t][3] + 5 for t in own], list(range(len(own))))
    for s, tid in enumerate(own):
        w = tl[tid][3] + 5
        assert np.array_equal(pool.download(s)[:w, :w], rec["tiles"][tid]), tid
