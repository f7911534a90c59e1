"""CPU tests of the ORACLE: pinned against the reference's own noise.cpp (compiled
unchanged into oracle/_ref when the reference checkout is present), the committed
known-answer vectors made from it (tests/golden/noise.json), the reference's data
fixture terrain4/DEM.dat (tests/golden/dem_dat.json) and the reference's own CPU
statement of the upsample filter (CPUElevationProducer)."""
import base64
import ctypes as C
import hashlib
import json
import os
import zlib

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
NOISE = json.load(open(os.path.join(HERE, "golden", "noise.json")))
DEM = json.load(open(os.path.join(HERE, "golden", "dem_dat.json")))
FRACTAL = [-140, -100, -15, -8, 5, 2.5, 1.5, 1, 0.5, 0.25, 0.1, 0.05]


def test_lcg_and_frandom_kats(oracle):
    seed = C.c_long(NOISE["lcg_seed"])
    for want_s, want_f in zip(NOISE["lcg"], NOISE["frandom_after_each"]):
        s2 = C.c_long(seed.value)
        assert oracle.lib().orc_frandom(C.byref(s2)) == np.float32(want_f)
        assert oracle.lib().orc_lrandom(C.byref(seed)) == want_s
    # seeds after 1..4 steps from 1234567
    assert NOISE["lcg"][:4] == [2026678708, 1074874845, 1238104402, 960255011]


def test_cnoise_kats(oracle):
    for (x, y), want in zip(NOISE["cnoise_points"], NOISE["cnoise"]):
        assert oracle.lib().orc_cnoise2(x, y) == np.float32(want), (x, y)


def test_cnoise_against_reference_build(oracle):
    ref = oracle.ref()
    if ref is None:
        pytest.skip("oracle/_ref not built (reference checkout absent)")
    rng = np.random.default_rng(1)
    pts = np.concatenate([rng.uniform(-3000, 9000, (20000, 2)),
                          rng.integers(-2000, 8000, (20000, 2)) + 0.5]).astype(np.float32)
    for x, y in pts:
        assert oracle.lib().orc_cnoise2(float(x), float(y)) == ref.ref_cnoise2(float(x), float(y))


def test_dem_noise_kats(oracle):
    n32 = oracle.dem_noise(101, r16f=False)
    n16 = oracle.dem_noise(101, r16f=True)
    # first centre values of layer 0 and their fp16 roundings
    # (printed to 8 significant digits there, hence the relative tolerance of one fp32 ulp)
    np.testing.assert_allclose(n32[0, 5, 5:9], [0.88749158, 0.00105512, 0.15307450, -0.10569286], rtol=2e-6, atol=0)
    np.testing.assert_allclose(n16[0, 5, 5:9], [0.88769531, 0.00105476, 0.15307617, -0.10571289], rtol=5e-6, atol=0)
    np.testing.assert_array_equal(n16, n32.astype(np.float16).astype(np.float32))   # RNE like numpy
    # first border values: seed 7654321 -> 0.05766881, seed 5647381 -> 0.30257297
    np.testing.assert_allclose(n32[0, 2, 5], 0.05766881, rtol=2e-6)    # layer 0: all borders seed A
    np.testing.assert_allclose(n32[5, 2, 5], 0.30257297, rtol=2e-6)    # layer 5: all borders seed B
    # corners stay zero, values in [-1, 1)
    assert np.all(n32[:, :5, :5] == 0) and np.all(n32[:, -5:, -5:] == 0)
    assert n32.min() >= -1 and n32.max() < 1


def test_dem_noise_borders_tile(oracle):
    """two tiles sharing an edge see the same noise along it: a layer's bottom border drawn with
    seed s equals the top border of a layer drawn with seed s, mirrored (ElevationProducer.cpp:50-128)"""
    n = oracle.dem_noise(101, r16f=False)
    W = 101
    # layer 0 (all seed A) must be symmetric under the border mirrorings it was built with
    a = n[0]
    np.testing.assert_array_equal(a[2, 5:W - 5], a[2, 5:W - 5][::-1])            # bottom row 2 symmetric
    np.testing.assert_array_equal(a[3, 5:W - 5], a[1, 5:W - 5][::-1])            # rows 3,4 mirrored into 1,0
    np.testing.assert_array_equal(a[4, 5:W - 5], a[0, 5:W - 5][::-1])
    # same seed on bottom and top: the top border is the bottom one shifted to the other side
    np.testing.assert_array_equal(a[W - 3, 5:W - 5], a[2, 5:W - 5])
    # layers differ only where their pattern bits differ: layer 1 (bit0 = bottom) vs layer 0
    b = n[1]
    assert not np.array_equal(a[0:5], b[0:5])
    np.testing.assert_array_equal(a[W - 5:, 5:W - 5], b[W - 5:, 5:W - 5])


def test_noise_select_tables(oracle):
    # every tile's 4 border bits must agree with its neighbours' (shared edges get the same bit):
    # right bit of (tx,ty) == left bit of (tx+1,ty), top bit == bottom bit of (tx,ty+1), face 0
    inv = {}
    rot = [0, 0, 1, 0, 2, 0, 1, 0, 3, 3, 1, 3, 2, 2, 1, 0]
    lay = [0, 1, 1, 2, 1, 3, 2, 4, 1, 2, 3, 4, 2, 4, 4, 5]
    level, n = 4, 16
    for ty in range(n):
        for tx in range(n):
            r, l = oracle.noise_select(level, tx, ty, 0)
            assert (r, l) in {(rot[b], lay[b]) for b in range(16)}
            inv[(tx, ty)] = (r, l)
    assert len(set(inv.values())) > 4


def test_texcoord_identities():
    """the shader's own fp32 coordinate maths resolves to the integer texels the oracle (and the
    kernels) use: parent taps, residual texel, noise texel and its rotations, zc lattice points
    (upsampleShader.glsl:141-186 with the uniforms of ElevationProducer.cpp:305-343)"""
    f = np.float32
    W, PW = 101, f(101)
    g = f(4)
    for dx in (0, 48):
        osl_x, osl_z = f(dx) / PW, f(1) / PW
        for x in range(W):
            st = f(x) + f(0.5)
            p_uv = np.floor(st) * f(0.5)
            off = (p_uv - (p_uv - np.floor(p_uv)) + f(0.5)) * osl_z + osl_x
            for i in range(4):
                tex = (off + f(i) * osl_z) * PW          # texel coordinate of the fetch
                assert int(np.floor(tex)) == x // 2 + dx + i
                assert abs(float(tex) - (x // 2 + dx + i + 0.5)) < 1e-3   # on the texel centre
            ij = np.floor(st - f(2))
            ux = f(2.5) + g * np.floor(ij / (f(2) * g) + f(0.5))
            vx = f(2.5) + g * np.floor(ij / (f(2) * g))
            assert int(np.floor((ux * osl_z + osl_x) * PW)) == 2 + 4 * ((x - 2 + 4) // 8) + dx
            assert int(np.floor((vx * osl_z + osl_x) * PW)) == 2 + 4 * ((x - 2) // 8) + dx
            ruv = p_uv * (f(2) / PW) + f(0.25) / PW           # residualOSH
            assert int(np.floor(ruv * PW)) == x
            nuv = (np.floor(st) + f(0.5)) / PW
            assert int(np.floor(nuv * PW)) == x and int(np.floor((f(1) - nuv) * PW)) == W - 1 - x


def test_residual_container_kats(oracle):
    hdr = DEM["header"]
    assert DEM["md5"] == "27c81f95bfa36e3ce67d238943041713" and DEM["size"] == 554338
    assert (hdr["minLevel"], hdr["maxLevel"], hdr["tileSize"], hdr["ntiles"], hdr["header_bytes"]) == (3, 7, 192, 344, 2780)
    prefix = base64.b64decode(DEM["prefix"])
    assert hashlib.sha1(prefix[28:]).hexdigest() == DEM["offsets_sha1"]
    # SURVEY 8c KATs (sha1 of the inflated int16 tiles, first 16 hex digits)
    kat = {0: "418364be53218510", 1: "978a9c5a3a05a5bd", 2: "6e49a9a428410bcc", 3: "597bcf5d18f77dfc",
           4: "a61d65a8d5120dc2", 100: "91f0a12ecbdf5161"}
    for t, h in kat.items():
        assert DEM["tile_sha1"][t].startswith(h)
    assert DEM["tile_width"][:4] == [29, 53, 101, 197]
    # the oracle's TIFF reader + zlib on the committed blobs
    for t, b64 in DEM["blobs"].items():
        blob = np.frombuffer(base64.b64decode(b64), np.uint8)
        w = DEM["tile_width"][int(t)]
        raw = np.empty(w * w * 2, np.uint8)
        ww, hh = C.c_int(), C.c_int()
        n = oracle.lib().orc_tiff_inflate(blob.ctypes.data_as(oracle.c_u8_p), C.c_uint32(len(blob)),
                                          raw.ctypes.data_as(oracle.c_u8_p), C.c_size_t(len(raw)),
                                          C.byref(ww), C.byref(hh))
        assert n == w * w * 2 and ww.value == w and hh.value == w
        assert hashlib.sha1(raw.tobytes()).hexdigest() == DEM["tile_sha1"][int(t)]


def test_residual_tile_ids(oracle):
    # a container with only the header + offsets is enough for the id / size arithmetic
    prefix = base64.b64decode(DEM["prefix"])
    r = oracle.Resid(prefix + b"\0" * 16, delta=2)
    assert [r.tile_size(l) for l in range(5)] == [24, 48, 96, 192, 192]
    assert [r.tile_id(l, 0, 0) for l in range(4)] == [0, 1, 2, 3]
    assert r.tile_id(4, 0, 0) == 4 and r.tile_id(4, 1, 1) == 7 and r.tile_id(5, 0, 0) == 8
    assert r.tile_id(7, 15, 15) == 343
    assert r.has_tile(0, 0, 0) and r.has_tile(5, 31, 31) and not r.has_tile(6, 0, 0)   # delta 2: levels 0..5


def test_upsample_filter_against_reference_cpu_statement(oracle):
    """orc_upsample_tile (GLSL restatement) vs CPUElevationProducer's own arithmetic on the same
    parent: same filter, different evaluation order -> equal within a few ulp of the heights."""
    rng = np.random.default_rng(3)
    W = 101
    parent = np.zeros((W, W, 3), np.float32)
    parent[:, :, 0] = rng.normal(0, 500, (W, W)).astype(np.float32)
    resid = rng.integers(-200, 200, (197, 197)).astype(np.float32)
    noise = np.zeros((6, W, W), np.float32)
    for (tx, ty) in ((0, 0), (1, 0), (0, 1), (3, 5)):
        p = oracle.elev_uniforms(3, tx, ty, noiseAmp=[0, 0, 0, 0], has_resid=1, resid_W=197)
        got = oracle.upsample_tile(p, parent, resid, noise)[:, :, 0]
        want = oracle.cpu_elevation_tile(W, 3, tx, ty, parent[:, :, 0], resid, 197, p.rx, p.ry)
        assert np.max(np.abs(got - want)) <= 4e-4          # |h| ~ 2000 -> ulp 1.2e-4


def test_canonical_vs_strict_tolerance(oracle):
    """The stated float tolerance: the canonical (fma) evaluation order the kernels share with the
    oracle versus the non-contracted reading of the same GLSL (liborc_strict.so), over a full
    quadtree to level 5: max |dh| <= 1e-5 of the height range, normals within 1 LSB of unorm8."""
    strict = C.CDLL(os.path.join(oracle.ORC_DIR, "liborc_strict.so"))
    W = 101
    noise = oracle.dem_noise(W)
    scene = oracle.make_scene(noiseAmp=FRACTAL, rootQuadSize=100000.0)

    def pair(lib, level, tx, ty, parent):
        e = np.empty((W, W, 3), np.float32)
        n = np.empty((W - 4, W - 4, 2), np.uint8)
        par = parent.ctypes.data_as(oracle.c_float_p) if parent is not None else None
        lib.orc_produce_pair(C.byref(scene), noise.ctypes.data_as(oracle.c_float_p), level, tx, ty, par, None,
                             e.ctypes.data_as(oracle.c_float_p), n.ctypes.data_as(oracle.c_u8_p))
        return e, n

    prev = {(0, 0): (None, None)}
    max_dh, max_dn, lo, hi = 0.0, 0, np.inf, -np.inf
    for level in range(6):
        cur = {}
        for ty in range(1 << level):
            for tx in range(1 << level):
                pa, pb = prev[(tx // 2, ty // 2)] if level else (None, None)
                ea, na = pair(oracle.lib(), level, tx, ty, pa)
                eb, nb = pair(strict, level, tx, ty, pb)
                cur[(tx, ty)] = (ea, eb)
                max_dh = max(max_dh, float(np.max(np.abs(ea - eb))))
                max_dn = max(max_dn, int(np.max(np.abs(na.astype(int) - nb.astype(int)))))
                lo, hi = min(lo, float(ea.min())), max(hi, float(ea.max()))
        prev = cur
    assert hi - lo > 100
    assert max_dh <= 1e-5 * (hi - lo), (max_dh, hi - lo)
    assert max_dn <= 1


def test_residual_builder_restatement_round_trips_through_the_reader():
    """orc_hm_encode_tile (HeightMipmap.cpp:449-559) against the reader it feeds: the approximation the
    builder carries equals upsample(parent) + residual * scale as ResidualProducer composes it
    (orc_resid_upsample + the int16 tile), and |heights - approximation| <= 0.5 (rounding)."""
    import numpy as np
    import orc
    rng = np.random.default_rng(7)
    n, ts = 29, 24                       # a 24 + 5 container whose level-1 tiles are full size
    yy, xx = np.mgrid[0:n, 0:n]
    parent = (300 * np.sin(xx / 5.0) * np.cos(yy / 7.0)).astype(np.float32)
    for tx, ty in ((0, 0), (1, 0), (0, 1), (1, 1)):
        tile = (parent[::1, ::1] * 0 + rng.normal(0, 40, (n, n))).astype(np.float32)
        resid, approx, mr, me = orc.hm_encode_tile(parent, tile, ts, tx, ty)
        assert resid.dtype == np.int16 and resid.shape == (ts + 5, ts + 5)
        assert me <= 0.5 + 1e-3 and mr >= np.abs(resid).max() - 0.5
        # the reader's side: upsample of the same parent quadrant, plus the stored integers
        f = orc.ResidFile()
        f.tileSize, f.minLevel = ts, 0
        up = np.zeros((n, n), np.float32)
        orc.lib().orc_resid_upsample(orc.C.byref(f), 1, tx, ty, parent.ctypes.data_as(orc.C.c_void_p),
                                     up.ctypes.data_as(orc.C.c_void_p))
        want = up[:ts + 5, :ts + 5] + resid.astype(np.float32)
        assert np.array_equal(approx[:ts + 5, :ts + 5], want)
