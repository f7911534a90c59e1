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


def test_noise_tables_are_kept_per_tile_width(plb, ctx, oracle):
    """demNoiseFactory caches one texture per tile width (ElevationProducer.cpp:135): initialising the noise of another
    width must not take the first producer's table away"""
    ctx.noise_init(101)
    other = ctx.noise_init(53)
    assert other.shape == (6, 53, 53)
    assert np.array_equal(ctx.noise_init(101), oracle.dem_noise(101).reshape(6, 101, 101))
    assert _run(plb, ctx, oracle, 2, noise_amp=FRACTAL) == 21          # still served from the 101-wide table


def test_pool_download_range_equals_per_slot_downloads(plb, ctx):
    rng = np.random.default_rng(5)
    for kind, w in ((plb.POOL_RESID_I16, 197), (plb.POOL_RESID_F32, 101), (plb.POOL_NORM2, 97), (plb.POOL_ELEV, 101)):
        pool = ctx.pool(kind, w, 7)
        shape, dt = pool._shape_dtype()
        for s in range(7):
            pool.upload(s, (rng.random(shape) * 200 - 50).astype(dt))
        got = pool.download_range(2, 4)
        for k in range(4):
            assert np.array_equal(got[k], pool.download(2 + k)), (kind, k)
        with pytest.raises(plb.PlError):
            pool.download_range(5, 3)


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


@pytest.mark.parametrize("sphere,elev_filter", [(1, 1), (1, 0), (0, 1), (0, 0)])
def test_fused_pair_kernel_equals_the_two_passes(plb, ctx, oracle, sphere, elev_filter):
    """pl_produce_range runs the fused elevation+normal kernel (pl_pair.cu); with pl_debug_no_fuse it
    launches the two passes separately.  Same tiles, same statistics, bit for bit -- and both equal
    the oracle (levels 0..4 of one face: negative, and at level 8+ amplitudes positive noise)."""
    amp = [-3250, -1590, 15, 8, 5]          # levels 2.. take the slope/curvature path
    kw = dict(noise_amp=amp, face=3 if sphere else 0, root_quad_size=12720000.0 if sphere else 100000.0,
              sphere=sphere, elev_filter=elev_filter)
    sc = plb.sweep_scene(want_stats=1, **kw)
    max_level = 4
    total = sum(4 ** l for l in range(max_level + 1))
    off = [sum(4 ** k for k in range(l)) for l in range(max_level + 1)]
    ctx.noise_init(101)
    out = []
    for fuse in (True, False):
        ctx.no_fuse(not fuse)
        elev = ctx.pool(plb.POOL_ELEV, 101, total)
        norm = ctx.pool(plb.POOL_NORM2, 97, total)
        n0 = ctx.launches
        for l in range(max_level + 1):
            ctx.produce_range(sc, elev, norm, l, 0, 4 ** l, off[l], off[l - 1] if l else 0, 0)
        ctx.sync()
        assert ctx.launches - n0 == (2 if fuse else 3) * (max_level + 1)
        out.append((elev, norm))
    ref = qt.oracle_quadtree(oracle, max_level, **kw)
    (fe, fn), (se, sn) = out
    for (l, tx, ty), (e, n, s) in ref.items():
        slot = off[l] + plb.morton_encode(tx, ty)
        a, b = fe.download(slot), se.download(slot)
        assert a.tobytes() == b.tobytes(), (l, tx, ty)
        assert np.array_equal(a, e), (l, tx, ty)
        a, b = fn.download(slot), sn.download(slot)
        assert a.tobytes() == b.tobytes(), (l, tx, ty)
        assert np.array_equal(a, n), (l, tx, ty)
    assert ctx.elev_stats_range(fe, 0, total).tobytes() == ctx.elev_stats_range(se, 0, total).tobytes()


def test_pair_batch_host_requests(plb, ctx, oracle):
    """pl_pair_batch: host-built request arrays for both passes, one fused launch per level; equals the
    oracle; a normal request that does not read the elevation tile of the same index is an argument
    error (nothing is launched)."""
    kw = dict(noise_amp=PLANET, face=5, root_quad_size=12720000.0, sphere=1)
    sc = plb.sweep_scene(want_stats=1, **kw)
    max_level = 3
    total = sum(4 ** l for l in range(max_level + 1))
    off = [sum(4 ** k for k in range(l)) for l in range(max_level + 1)]
    elev = ctx.pool(plb.POOL_ELEV, 101, total)
    norm = ctx.pool(plb.POOL_NORM2, 97, total)
    ctx.noise_init(101)
    n0 = ctx.launches
    for l in range(max_level + 1):
        e, q = plb.make_requests_range(sc, l, 0, 4 ** l, off[l], off[l - 1] if l else 0, 0)
        ctx.pair_batch(sc.elev, sc.norm, elev, norm, e, q)
    ctx.sync()
    assert ctx.launches - n0 == max_level + 1
    ref = qt.oracle_quadtree(oracle, max_level, **kw)
    for (l, tx, ty), (e, n, s) in ref.items():
        slot = off[l] + plb.morton_encode(tx, ty)
        assert np.array_equal(elev.download(slot), e), (l, tx, ty)
        assert np.array_equal(norm.download(slot), n), (l, tx, ty)
    e, q = plb.make_requests_range(sc, 1, 0, 4, off[1], 0, 0)
    q["elev_slot"][2] = off[1]          # reads another tile's elevation
    with pytest.raises(plb.PlError) as err:
        ctx.pair_batch(sc.elev, sc.norm, elev, norm, e, q)
    assert err.value.code == plb.PL_ERR_ARG
    assert ctx.launches - n0 == max_level + 1


def test_async_stats_readback(plb, ctx):
    """pl_elev_stats_readback_begin/_end deliver the values pl_elev_stats_range does, also when more
    work was enqueued (and the slots rewritten) between begin and end; at most 4 are in flight."""
    sc = plb.sweep_scene(want_stats=1, noise_amp=FRACTAL)
    elev = ctx.pool(plb.POOL_ELEV, 101, 21)
    norm = ctx.pool(plb.POOL_NORM2, 97, 21)
    ctx.noise_init(101)
    off = [0, 1, 5]
    for l in range(3):
        ctx.produce_range(sc, elev, norm, l, 0, 4 ** l, off[l], off[l - 1] if l else 0, 0)
    want = ctx.elev_stats_range(elev, 5, 16)
    tk = ctx.elev_stats_readback_begin(elev, 5, 16)
    sc2 = plb.sweep_scene(want_stats=1, noise_amp=[x * 2 for x in FRACTAL])
    ctx.produce_range(sc2, elev, norm, 2, 0, 16, 5, 1, 0)          # rewrites slots 5..20 behind the read-back
    got = ctx.elev_stats_readback_end(tk)
    assert got.tobytes() == want.tobytes()
    assert ctx.elev_stats_range(elev, 5, 16).tobytes() != want.tobytes()
    tks = [ctx.elev_stats_readback_begin(elev, 0, 1) for _ in range(4)]
    with pytest.raises(plb.PlError):
        ctx.elev_stats_readback_begin(elev, 0, 1)
    for t in tks:
        ctx.elev_stats_readback_end(t)
    with pytest.raises(plb.PlError):
        ctx.elev_stats_readback_end(tks[0])


def test_request_staging_ring_wraps_and_grows(plb, ctx, oracle):
    """host request arrays travel through a FIFO staging ring (pinned + device halves, uploads on a copy
    stream).  With a 256 KB ring the level-4 batches (78 KB) wrap it every third batch while earlier
    uploads and kernels are still in flight, and the level-5 batch (311 KB) makes it grow; the tiles
    must come out as if every batch had been staged on its own."""
    ctx.stage_ring(256 << 10)
    kw = dict(noise_amp=PLANET, face=4, root_quad_size=12720000.0, sphere=1)
    sc = plb.sweep_scene(want_stats=1, **kw)
    max_level = 5
    total = sum(4 ** l for l in range(max_level + 1))
    off = [sum(4 ** k for k in range(l)) for l in range(max_level + 1)]
    elev = ctx.pool(plb.POOL_ELEV, 101, total)
    norm = ctx.pool(plb.POOL_NORM2, 97, total)
    ctx.noise_init(101)
    reqs = [plb.make_requests_range(sc, l, 0, 4 ** l, off[l], off[l - 1] if l else 0, 0) for l in range(max_level + 1)]
    for rep in range(12):                       # levels 0..4 over and over: 60 batches, ~1.2 MB through 256 KB
        for l in range(5):
            ctx.pair_batch(sc.elev, sc.norm, elev, norm, *reqs[l])
    ctx.pair_batch(sc.elev, sc.norm, elev, norm, *reqs[5])        # grows the ring with batches in flight
    for l in range(5):
        ctx.elevation_batch(sc.elev, elev, reqs[l][0])            # the single-array users share the ring
        ctx.normal_batch(sc.norm, norm, elev, reqs[l][1])
    ctx.sync()
    ref = qt.oracle_quadtree(oracle, 4, **kw)
    for (l, tx, ty), (e, n, s) in ref.items():
        slot = off[l] + plb.morton_encode(tx, ty)
        assert np.array_equal(elev.download(slot), e), (l, tx, ty)
        assert np.array_equal(norm.download(slot), n), (l, tx, ty)
    # level 5 against the device-generated path
    elev2 = ctx.pool(plb.POOL_ELEV, 101, total)
    norm2 = ctx.pool(plb.POOL_NORM2, 97, total)
    for l in range(max_level + 1):
        ctx.produce_range(sc, elev2, norm2, l, 0, 4 ** l, off[l], off[l - 1] if l else 0, 0)
    for slot in range(off[5], total, 37):
        assert elev.download(slot).tobytes() == elev2.download(slot).tobytes(), slot
        assert norm.download(slot).tobytes() == norm2.download(slot).tobytes(), slot


# ----------------------------------------------------------------- residuals

import base64
import hashlib
import json
import os
import zlib

import resid_synth as rs

DEM = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dem_dat.json")))


@pytest.fixture(params=[1, 2], ids=["warp-per-stream", "tokenizer+resolver"])
def inflate_path(request, ctx):
    """both DEFLATE decoders of pl_residual.cu (the library picks one by batch size: pl_debug_inflate_path)"""
    ctx.inflate_path(request.param)
    yield request.param
    ctx.inflate_path(0)


def test_residual_decode_all_344_tiles_of_the_reference_fixture(plb, ctx, inflate_path):
    """terrain4/DEM.dat whole (tests/golden/terrain4_DEM.dat, md5 in dem_dat.json): every one of its 344 tiles through the
    device decoder, sha1 of the inflated int16 bytes against the reference file inflated by zlib (make_golden.py)"""
    import struct
    data = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "terrain4_DEM.dat"), "rb").read()
    assert hashlib.md5(data).hexdigest() == DEM["md5"]
    nt, header = DEM["header"]["ntiles"], DEM["header"]["header_bytes"]
    offs = struct.unpack_from("<%dI" % (2 * nt), data, 28)
    blobs = [data[header + offs[2 * t]: header + offs[2 * t + 1]] for t in range(nt)]
    widths = DEM["tile_width"]
    pool = ctx.pool(plb.POOL_RESID_I16, 197, nt)
    ctx.residual_decode(pool, blobs, widths, list(range(nt)))
    for t, w in enumerate(widths):
        got = pool.download(t)[:w, :w]
        assert hashlib.sha1(np.ascontiguousarray(got, "<i2").tobytes()).hexdigest() == DEM["tile_sha1"][t], t
    # the same file as an archive resident in device memory (pl_blobs_create / pl_residual_decode_stored): tiles located
    # by the file's own offset table, decoded in place
    store = ctx.blobs(data)
    pool2 = ctx.pool(plb.POOL_RESID_I16, 197, nt)
    ctx.residual_decode_stored(pool2, store, [header + offs[2 * t] for t in range(nt)],
                               [offs[2 * t + 1] - offs[2 * t] for t in range(nt)], widths, list(range(nt)))
    for t in range(nt):
        assert np.array_equal(pool2.download(t), pool.download(t)), t
    with pytest.raises(plb.PlError):
        ctx.residual_decode_stored(pool2, store, [len(data) - 10], [100], [197], [0])      # outside the archive
    store.close()


def test_residual_decode_reference_fixture(plb, ctx, inflate_path):
    """the reference's own fixture (terrain4/DEM.dat): inflated int16 tiles, sha1 by sha1"""
    ids = sorted(int(t) for t in DEM["blobs"])
    blobs = [base64.b64decode(DEM["blobs"][str(t)]) for t in ids]
    widths = [DEM["tile_width"][t] for t in ids]
    pool = ctx.pool(plb.POOL_RESID_I16, 197, len(ids))
    ctx.residual_decode(pool, blobs, widths, list(range(len(ids))))
    for s, (t, w) in enumerate(zip(ids, widths)):
        got = pool.download(s)[:w, :w]
        assert hashlib.sha1(np.ascontiguousarray(got, "<i2").tobytes()).hexdigest() == DEM["tile_sha1"][t], t
    # float pool: (float) z * scale, lower-left corner of the 197-stride slot, rest untouched
    fpool = ctx.pool(plb.POOL_RESID_F32, 197, len(ids) + 1)
    ctx.residual_decode(fpool, blobs, widths, list(range(len(ids))), scale=2.5)
    for s, w in enumerate(widths):
        f = fpool.download(s)
        np.testing.assert_array_equal(f[:w, :w], pool.download(s)[:w, :w].astype(np.float32) * np.float32(2.5))
        assert np.all(f[w:, :] == 0) and np.all(f[:, w:] == 0)


@pytest.mark.parametrize("level,strategy", [(0, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_DEFAULT_STRATEGY),
                                            (6, zlib.Z_DEFAULT_STRATEGY), (9, zlib.Z_FILTERED),
                                            (6, zlib.Z_FIXED), (6, zlib.Z_RLE), (6, zlib.Z_HUFFMAN_ONLY)])
def test_residual_decode_deflate_variants(plb, ctx, level, strategy, inflate_path):
    """stored, fixed-Huffman and dynamic-Huffman blocks; long matches (constant tiles), incompressible noise"""
    rng = np.random.default_rng(level * 10 + strategy)
    tiles = [rs.fractal_tile(rng, 197, 300), np.zeros((197, 197), np.int16),
             rng.integers(-32768, 32767, (197, 197)).astype(np.int16), rs.fractal_tile(rng, 53, 40),
             np.full((29, 29), -7, np.int16), (np.arange(197 * 197).reshape(197, 197) % 251 - 100).astype(np.int16)]
    blobs = [rs.tiff_blob(t, level, strategy) for t in tiles]
    blobs.append(rs.tiff_blob(tiles[0], compression=1))       # uncompressed strip
    tiles.append(tiles[0])
    pool = ctx.pool(plb.POOL_RESID_I16, 197, len(tiles))
    ctx.residual_decode(pool, blobs, [t.shape[0] for t in tiles], list(range(len(tiles))))
    for s, t in enumerate(tiles):
        np.testing.assert_array_equal(pool.download(s)[:t.shape[0], :t.shape[0]], t)


def test_residual_decode_root_composition_and_errors(plb, ctx, inflate_path):
    rng = np.random.default_rng(11)
    a, b = rs.fractal_tile(rng, 101, 50), rs.fractal_tile(rng, 101, 10)
    pool = ctx.pool(plb.POOL_RESID_F32, 197, 3)
    base = rng.normal(0, 100, (197, 197)).astype(np.float32)
    pool.upload(1, base)
    # result = tile[add_slot] + int16 * scale (ResidualProducer.cpp:333-336), in place allowed
    ctx.residual_decode(pool, [rs.tiff_blob(a), rs.tiff_blob(b)], [101, 101], [0, 1], add_slots=[-1, 1], scale=0.5)
    np.testing.assert_array_equal(pool.download(0)[:101, :101], a.astype(np.float32) * np.float32(0.5))
    np.testing.assert_array_equal(pool.download(1)[:101, :101], base[:101, :101] + b.astype(np.float32) * np.float32(0.5))
    np.testing.assert_array_equal(pool.download(1)[101:, :], base[101:, :])
    good = rs.tiff_blob(a)
    bad = bytearray(good)
    bad[40] ^= 0x5A                                            # corrupt the DEFLATE stream
    with pytest.raises(plb.PlError) as e:
        ctx.residual_decode(pool, [bytes(bad)], [101], [2])
    assert e.value.code == plb.PL_ERR_CORRUPT
    with pytest.raises(plb.PlError) as e:
        ctx.residual_decode(pool, [good[:30]], [101], [2])     # truncated: not a TIFF
    assert e.value.code == plb.PL_ERR_CORRUPT
    with pytest.raises(plb.PlError) as e:
        ctx.residual_decode(pool, [good], [53], [2])           # width mismatch
    assert e.value.code == plb.PL_ERR_CORRUPT


def test_residual_decode_rejects_what_zlib_rejects(plb, ctx, inflate_path):
    """an INCOMPLETE literal/length code is an error in zlib (inftrees.c: "invalid literal/lengths set"), i.e. in the
    reference's TIFFReadEncodedStrip; the same hand-written stream with a complete code decodes.  Both decoders."""
    import zlib
    good, bad = rs.hand_made_zlib_stream(True), rs.hand_made_zlib_stream(False)
    assert zlib.decompress(good) == b"AA"
    with pytest.raises(zlib.error):
        zlib.decompress(bad)
    one = np.zeros((1, 1), np.int16)
    pool = ctx.pool(plb.POOL_RESID_I16, 197, 1)
    ctx.residual_decode(pool, [rs.tiff_blob(one, strip=good)], [1], [0])
    assert pool.download(0)[0, 0] == 0x4141
    with pytest.raises(plb.PlError) as e:
        ctx.residual_decode(pool, [rs.tiff_blob(one, strip=bad)], [1], [0])
    assert e.value.code == plb.PL_ERR_CORRUPT
    # the Adler-32 trailer (RFC 1950) is verified like inflate() does: a structurally valid stream whose checksum is
    # wrong, or missing, fails
    for strip in (good[:-1] + bytes([good[-1] ^ 1]), good[:-4]):
        with pytest.raises(zlib.error):
            zlib.decompress(strip)
        with pytest.raises(plb.PlError) as e:
            ctx.residual_decode(pool, [rs.tiff_blob(one, strip=strip)], [1], [0])
        assert e.value.code == plb.PL_ERR_CORRUPT
    rng = np.random.default_rng(3)
    tile = rs.fractal_tile(rng, 197, 80)
    co = zlib.compressobj(6)
    strip = bytearray(co.compress(tile.tobytes()) + co.flush())
    ctx.residual_decode(pool, [rs.tiff_blob(tile, strip=bytes(strip))], [197], [0])
    assert np.array_equal(pool.download(0), tile)
    strip[-2] ^= 0x10
    with pytest.raises(plb.PlError) as e:
        ctx.residual_decode(pool, [rs.tiff_blob(tile, strip=bytes(strip))], [197], [0])
    assert e.value.code == plb.PL_ERR_CORRUPT


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("kind", ["f32", "i16"])
def test_elevation_with_residuals(plb, ctx, oracle, kind, fused):
    """(fused: both passes through pl_pair_batch, the fused kernel's residual variants)
    config 3 shape: 197-wide residual tiles (mod 2), flip, NEAREST storage, sphere normals;
    residual float tiles from the oracle's ResidualProducer restatement (delta = 0) uploaded to
    the F32 pool, or the raw int16 tiles consumed directly from an I16 pool"""
    data, tiles = rs.container(min_level=0, max_level=2, tile_size=192, scale=0.5, zero_fraction=0.2)
    res = oracle.Resid(data, delta=0)
    amp = [0, 0, 5, 2.5, 1]
    W = 101
    kw = dict(noise_amp=amp, face=2, root_quad_size=12720000.0, sphere=1, flip=1, elev_filter=0)
    scene = oracle.make_scene(W=W, rootQuadSize=kw["root_quad_size"], face=2, flip=1, noiseAmp=amp, sphere=1,
                              elev_filter=0, resid=res)
    noise = oracle.dem_noise(W)
    max_level = 3
    ntile = sum(4 ** l for l in range(max_level + 1))
    elev = ctx.pool(plb.POOL_ELEV, W, ntile)
    norm = ctx.pool(plb.POOL_NORM2, W - 4, ntile)
    rpool = ctx.pool(plb.POOL_RESID_F32 if kind == "f32" else plb.POOL_RESID_I16, 197, 32)
    ctx.noise_init(W)
    es = plb.elev_scene(W, 24, 1, 1, 0, 1, resid_scale=0.5)
    ns = plb.norm_scene(W - 4, 24, 2, 0, 1, 1)
    # residual tiles: one per (level, tx/2, ty/2) that the file has
    rslot = {}
    for l in range(max_level + 1):
        for ty in range(max(1, (1 << l) // 2)):
            for tx in range(max(1, (1 << l) // 2)):
                if res.has_tile(l, tx, ty):
                    s = len(rslot)
                    rslot[(l, tx, ty)] = s
                    if kind == "f32":
                        rpool.upload(s, res.create_tile(l, tx, ty))
                    else:
                        rpool.upload(s, tiles[rs.tile_id(0, l, tx, ty)])
    slot_of, ref = {}, {}
    for l in range(max_level + 1):
        tl = qt.level_tiles(l)
        has = [res.has_tile(l, tx // 2, ty // 2) for (_, tx, ty) in tl]
        reqs = plb.elev_make_reqs(tl, tile_w=W, root_quad_size=kw["root_quad_size"], noise_amp=amp, face=2,
                                  resid_tile_w=197, has_resid=has)
        nreqs = plb.norm_make_reqs(tl, ns, root_quad_size=kw["root_quad_size"])
        for i, t in enumerate(tl):
            slot_of[t] = len(slot_of)
            reqs["out_slot"][i] = nreqs["out_slot"][i] = nreqs["elev_slot"][i] = slot_of[t]
            if l:
                reqs["parent_slot"][i] = slot_of[(l - 1, t[1] // 2, t[2] // 2)]
            if has[i]:
                reqs["resid_slot"][i] = rslot[(l, t[1] // 2, t[2] // 2)]
            parent = ref[(l - 1, t[1] // 2, t[2] // 2)][0] if l else None
            rt = res.create_tile(l, t[1] // 2, t[2] // 2) if has[i] else None
            ref[t] = oracle.produce_pair(scene, noise, l, t[1], t[2], parent, rt)
        if fused:
            ctx.pair_batch(es, ns, elev, norm, reqs, nreqs, resid=rpool)
        else:
            ctx.elevation_batch(es, elev, reqs, resid=rpool)
            ctx.normal_batch(ns, norm, elev, nreqs)
    assert any(res.has_tile(3, x, y) for x in range(4) for y in range(4)) is False   # maxLevel 2: level 3 is noise only
    for t, (e, n) in ref.items():
        assert np.array_equal(elev.download(slot_of[t]), e), t
        assert np.array_equal(norm.download(slot_of[t]), n), t
    # host arrays are validated: a window outside the residual / parent tile, a misaligned one, a bad residual slot
    l2 = [i for i in range(len(reqs)) if reqs["resid_slot"][i] >= 0] or [0]
    for field, value in (("rx", 100), ("rx", 2), ("ry", -1), ("ry", 97), ("dx", 3), ("dy", 96), ("resid_slot", -2), ("parent_slot", -5)):
        bad = reqs.copy()
        bad["resid_slot"][l2[0]] = max(int(bad["resid_slot"][l2[0]]), 0)
        bad[field][l2[0]] = value
        with pytest.raises(plb.PlError):
            if fused:
                ctx.pair_batch(es, ns, elev, norm, bad, nreqs, resid=rpool)
            else:
                ctx.elevation_batch(es, elev, bad, resid=rpool)


@pytest.mark.parametrize("path", ["two passes", "fused", "produce_range"])
@pytest.mark.parametrize("sphere,parent_filter", [(0, 1), (1, 1), (1, 0)])
def test_rgba8_normals_with_parent_coarse_normal(plb, ctx, oracle, sphere, parent_filter, path):
    """4-channel normal storages (tileSDF.z = 1, normalShader.glsl:100-119): .zw = the parent
    tile's normal at the coarse mesh vertices, rotated by parentToTangentFrame on a sphere.  Through the runtime-geometry
    normal kernel (two passes), the fused elevation + normal kernel (pl_pair_batch: its RGBA8 variant) and
    pl_produce_range (device-generated requests naming the parent's normal tile)."""
    W, rq = 101, 12720000.0 if sphere else 100000.0
    amp = PLANET if sphere else FRACTAL
    tiles = [t for l in range(3) for t in qt.level_tiles(l)]
    slot = {t: i for i, t in enumerate(tiles)}
    elev = ctx.pool(plb.POOL_ELEV, W, len(tiles))
    norm = ctx.pool(plb.POOL_NORM4, W - 4, len(tiles))
    ctx.noise_init(W)
    es = plb.elev_scene(W, 24, 0, 1, 0, 0)
    ns = plb.norm_scene(W - 4, 24, 2, 1, parent_filter, sphere)
    for level in range(3):
        lt = qt.level_tiles(level)
        reqs = plb.elev_make_reqs(lt, tile_w=W, root_quad_size=rq, noise_amp=amp, face=3 if sphere else 0)
        nreqs = plb.norm_make_reqs(lt, ns, root_quad_size=rq, components=4)
        for i, t in enumerate(lt):
            reqs["out_slot"][i] = nreqs["out_slot"][i] = nreqs["elev_slot"][i] = slot[t]
            if level > 0:
                reqs["parent_slot"][i] = nreqs["parent_slot"][i] = slot[(level - 1, t[1] // 2, t[2] // 2)]
        if path == "two passes":
            ctx.elevation_batch(es, elev, reqs)
            ctx.normal_batch(ns, norm, elev, nreqs)
        elif path == "fused":
            ctx.pair_batch(es, ns, elev, norm, reqs, nreqs)
        else:     # slots of qt.level_tiles are row-major per level; produce_range numbers them in Morton order
            sc = plb.sweep_scene(noise_amp=amp, face=3 if sphere else 0, root_quad_size=rq, sphere=sphere, want_stats=0)
            sc.norm.parent_filter = parent_filter
            off = [0, 1, 5]
            ctx.produce_range(sc, elev, norm, level, 0, 4 ** level, off[level], off[level - 1] if level else 0, 0)
    ctx.sync()
    if path == "produce_range":
        off = [0, 1, 5]
        slot = {(l, tx, ty): off[l] + plb.morton_encode(tx, ty) for (l, tx, ty) in tiles}
    ref = {}
    for t in tiles:
        level, tx, ty = t
        e = elev.download(slot[t])
        p = oracle.normal_uniforms(level, tx, ty, components=4, parent_filter=parent_filter, rootQuadSize=rq, sphere=sphere)
        parent = None
        if level > 0:   # what the GL sampler returns for the parent's RGBA8 texels: c / 255
            parent = ref[(level - 1, tx // 2, ty // 2)].astype(np.float32) / np.float32(255.0)
        ref[t] = oracle.pack_unorm8(oracle.normal_tile(p, e, parent), 4)
        got = norm.download(slot[t])
        assert np.array_equal(got, ref[t]), "tile %r: %d bytes differ" % (t, np.count_nonzero(got != ref[t]))
    # the coarse channels really differ from the fine ones below the root
    deep = ref[(2, 1, 2)]
    assert np.count_nonzero(deep[..., 2:] != deep[..., :2]) > 1000


@pytest.mark.parametrize("delta", [1, 2, 3])
def test_residual_root_composition_delta(plb, ctx, oracle, delta):
    """ResidualProducer.cpp:218-228 + upsample (:342-384): with delta > 0 the root tile is the stored
    level-0 tile upsampled `delta` times, each time adding the stored residual of that level
    (earth-srtm.xml: delta="2"; DEM.dat geometry: minLevel 3, tileSize 192 -> 29, 53, 101, 197)"""
    data, _ = rs.container(min_level=3, max_level=4, tile_size=192, scale=0.25, seed=77 + delta)
    res = oracle.Resid(data, delta=delta)
    want = res.create_tile(0, 0, 0)                      # stored level `delta`, composed
    pool = ctx.pool(plb.POOL_RESID_F32, 197, 2)
    pool.upload(0, np.full((197, 197), 7.0, np.float32))   # stale content must not leak in
    ctx.residual_decode(pool, [res.blob(0)], [res.tile_size(0) + 5], [0], scale=0.25)
    for i in range(1, delta + 1):
        ts = res.tile_size(i)
        ctx.residual_upsample(pool, 0, plb.SLOT_SCRATCH, ts)
        ctx.residual_decode(pool, [res.blob(res.tile_id(i, 0, 0))], [ts + 5], [0], add_slots=[plb.SLOT_SCRATCH], scale=0.25)
    w = res.tile_size(delta) + 5
    got = pool.download(0)
    assert np.array_equal(got[:w, :w], want[:w, :w])
    assert np.abs(want[:w, :w]).max() > 10
    with pytest.raises(plb.PlError):
        ctx.residual_upsample(pool, 0, 0, 96)            # in place is not defined
    with pytest.raises(plb.PlError):
        ctx.residual_upsample(pool, 0, 1, 194)           # does not fit the pool


# ------------------------------------------------- the step before the path: building residual files

def _height_levels(rng, top=24, levels=5):
    """int16 height fields of a square domain, level l = top * 2^l samples (+1), coarser levels are
    point samples of the next finer one (HeightMipmap::buildMipmapLevel, HeightMipmap.cpp:199-253)"""
    n = top << (levels - 1)
    yy, xx = np.mgrid[0:n + 1, 0:n + 1] / n
    base = 900 * np.sin(5 * xx) * np.cos(4 * yy) + 250 * np.sin(23 * xx + 2) * np.sin(17 * yy) + rng.normal(0, 6, xx.shape)
    out = [np.rint(base).astype(np.int16)]
    for _ in range(levels - 1):
        out.insert(0, out[0][::2, ::2])
    return out


def _height_tile(field, ts, tx, ty, n):
    """the (ts + 5)^2 tile with its 2-sample border, edge samples repeated, in an n x n float array"""
    idx_x = np.clip(np.arange(-2, ts + 3) + tx * ts, 0, field.shape[1] - 1)
    idx_y = np.clip(np.arange(-2, ts + 3) + ty * ts, 0, field.shape[0] - 1)
    out = np.zeros((n, n), np.float32)
    out[:ts + 5, :ts + 5] = field[np.ix_(idx_y, idx_x)].astype(np.float32)
    return out


def test_residual_pyramid_builder_and_round_trip(plb, ctx, oracle):
    """pl_residual_encode_batch (HeightMipmap::buildResiduals on the device) against the oracle, level by
    level, then the loop closed: the residuals go into a container in the reference's format, come back
    through pl_residual_decode_batch, and ResidualProducer's composition upsample(parent) + residual
    reproduces the builder's approximation of every tile bit for bit."""
    rng = np.random.default_rng(11)
    min_level, max_level, tile_size, n = 3, 4, 192, 197
    fields = _height_levels(rng, 24, max_level + 1)
    tiles_of = lambda l: 1 if l < min_level else 1 << (l - min_level)
    ts_of = lambda l: rs.tile_width(min_level, tile_size, l) - 5
    ids = {(l, tx, ty): rs.tile_id(min_level, l, tx, ty) for l in range(max_level + 1)
           for ty in range(tiles_of(l)) for tx in range(tiles_of(l))}
    nt = len(ids)
    heights = ctx.pool(plb.POOL_RESID_F32, n, nt)
    approx = ctx.pool(plb.POOL_RESID_F32, n, nt + 1)
    resid = ctx.pool(plb.POOL_RESID_I16, n, nt)
    want_approx, resid_tiles = {}, {}
    for key, tid in ids.items():
        heights.upload(tid, _height_tile(fields[key[0]], ts_of(key[0]), key[1], key[2], n))
    # level 0: stored as is, and its own approximation (produceTile :567-578, getApproxTile :420-431)
    t0 = _height_tile(fields[0], ts_of(0), 0, 0, n)
    approx.upload(0, t0)
    want_approx[(0, 0, 0)] = t0
    resid_tiles[0] = np.rint(t0[:ts_of(0) + 5, :ts_of(0) + 5]).astype(np.int16)
    for l in range(1, max_level + 1):
        keys = [k for k in ids if k[0] == l]
        reqs = np.zeros(len(keys), plb.RESID_ENC_DTYPE)
        for i, (_, tx, ty) in enumerate(keys):
            parent = (l - 1, tx // 2, ty // 2) if l > min_level else (l - 1, 0, 0)
            reqs[i] = (ids[(l, tx, ty)], ids[parent], ids[(l, tx, ty)], ids[(l, tx, ty)], ts_of(l), tx, ty, 0)
        mr, me = ctx.residual_encode(heights, approx, resid, reqs)
        for i, key in enumerate(keys):
            _, tx, ty = key
            parent = (l - 1, tx // 2, ty // 2) if l > min_level else (l - 1, 0, 0)
            r, a, omr, ome = oracle.hm_encode_tile(want_approx[parent], _height_tile(fields[l], ts_of(l), tx, ty, n),
                                                   ts_of(l), tx, ty)
            w = ts_of(l) + 5
            assert np.array_equal(resid.download(ids[key])[:w, :w], r), key
            assert np.array_equal(approx.download(ids[key])[:w, :w], a[:w, :w]), key
            assert mr[i] == np.float32(omr) and me[i] == np.float32(ome) and ome <= 0.5, key
            want_approx[key] = a
            resid_tiles[ids[key]] = r
    # the loop closed: container -> decode -> upsample(parent approximation) + residual == approximation
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        plb.residual_write_file(os.path.join(tmp, "DEM.dat"), resid_tiles, min_level, max_level, tile_size)
        data = open(os.path.join(tmp, "DEM.dat"), "rb").read()
    assert oracle.Resid(rs.container_from_tiles(resid_tiles, min_level, max_level, tile_size)).inflate(5)[0] == \
        oracle.Resid(data).inflate(5)[0]
    rd = oracle.Resid(data)
    dec = ctx.pool(plb.POOL_RESID_F32, n, nt + 1)
    dec.upload(0, want_approx[(0, 0, 0)])
    for l in range(1, max_level + 1):
        for key in [k for k in ids if k[0] == l]:
            _, tx, ty = key
            parent = (l - 1, tx // 2, ty // 2) if l > min_level else (l - 1, 0, 0)
            ctx.residual_upsample(dec, ids[parent], nt, ts_of(l), tx, ty)           # into the spare slot
            ctx.residual_decode(dec, [rd.blob(ids[key])], [ts_of(l) + 5], [ids[key]], add_slots=[nt], scale=1.0)
            w = ts_of(l) + 5
            assert np.array_equal(dec.download(ids[key])[:w, :w], want_approx[key][:w, :w]), key
    # and the elevation the run-time path would start from stays within the quantisation error
    fine = want_approx[(max_level, 1, 1)][2:195, 2:195]
    truth = fields[max_level][192:385, 192:385].astype(np.float32)
    assert np.abs(fine - truth).max() <= 0.5
