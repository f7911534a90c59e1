"""CPU tests of the HOST side of the product library: the C ABI surface, the per-tile uniform
maths (against the oracle), the sweep planner, and the no-GPU error behaviour."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLANET = [-3250, -1590, -1125, -795, -561, -397, -140, -100, 15, 8, 5, 2.5, 1.5, 1, 0.5, 0.25, 0.1, 0.05]


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol(plb):
    hdr = open(os.path.join(ROOT, "include", "proland_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(pl_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(plb.EXPORTS), declared ^ set(plb.EXPORTS)
    lib = plb.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), "libproland_b200.so does not export " + name
    assert lib.pl_abi_version() == 2


def test_no_device_is_an_error_not_a_fallback(plb):
    if _has_gpu():
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    rc = plb.lib().pl_ctx_create(0, C.byref(h))
    assert rc == plb.PL_ERR_NO_DEVICE and not h.value
    assert b"no CPU fallback" in plb.lib().pl_last_error()
    with pytest.raises(plb.PlError):
        plb.Context(0)


def test_cnoise_matches_oracle(plb, oracle):
    rng = np.random.default_rng(5)
    pts = np.concatenate([rng.uniform(-5000, 9000, (30000, 2)),
                          rng.integers(-5000, 9000, (30000, 2)) + 0.5]).astype(np.float32)
    for x, y in pts:
        assert plb.cnoise2(float(x), float(y)) == oracle.lib().orc_cnoise2(float(x), float(y))


@pytest.mark.parametrize("face", [0, 1, 2, 3, 4, 5, 6])
def test_noise_select_matches_oracle(plb, oracle, face):
    for level in (0, 1, 2, 5):
        n = 1 << level
        for ty in range(n):
            for tx in range(n):
                assert plb.noise_select(level, tx, ty, face) == oracle.noise_select(level, tx, ty, face)
    for level in (9, 12, 14):      # edges and a diagonal of deep levels (negative lattice coordinates on face 6)
        n = 1 << level
        for k in range(0, n, max(1, n // 97)):
            for tx, ty in ((k, 0), (0, k), (n - 1, k), (k, n - 1), (k, k)):
                assert plb.noise_select(level, tx, ty, face) == oracle.noise_select(level, tx, ty, face)


@pytest.mark.parametrize("face,sphere,size", [(0, 0, 100000.0), (3, 1, 12720000.0), (6, 1, 12720000.0)])
def test_requests_match_oracle_uniforms(plb, oracle, face, sphere, size):
    sc = plb.sweep_scene(noise_amp=PLANET, face=face, root_quad_size=size, sphere=sphere)
    for level in (0, 1, 3, 6, 10):
        n = min(4 ** level, 1024)
        m0 = (4 ** level - n) // 2 // 4 * 4
        e, q = plb.make_requests_range(sc, level, m0, n, 100, 7, m0 >> 2, nthreads=3)
        for i in range(0, n, max(1, n // 64)):
            tx, ty = plb.morton_decode(m0 + i)
            p = oracle.elev_uniforms(level, tx, ty, rootQuadSize=size, noiseAmp=PLANET, face=face)
            assert (e["dx"][i], e["dy"][i], e["noise_r"][i], e["noise_l"][i]) == (p.dx, p.dy, p.noiseR, p.noiseL)
            assert e["rs"][i] == np.float32(p.rs) and e["pixel_size"][i] == np.float32(p.pixel_size)
            assert (e["level"][i], e["tx"][i], e["ty"][i]) == (level, tx, ty)
            assert e["out_slot"][i] == 100 + i
            assert e["parent_slot"][i] == (7 + ((m0 + i) >> 2) - (m0 >> 2) if level else -1)
            u = oracle.normal_uniforms(level, tx, ty, rootQuadSize=size, sphere=sphere)
            np.testing.assert_array_equal(q["deform"][i], np.float32(u.deform[:]))
            np.testing.assert_array_equal(q["w2t"][i], np.float32(u.w2t[:]))
            if sphere:
                np.testing.assert_array_equal(q["corners"][i], np.float32(u.corners[:12]))
                np.testing.assert_array_equal(q["verticals"][i], np.float32(u.verticals[:12]))
                np.testing.assert_array_equal(q["norms"][i], np.float32(u.norms[:]))
                if level > 0:
                    np.testing.assert_array_equal(q["p2t"][i], np.float32(u.p2t[:]))


def test_single_request_entry_points(plb):
    sc = plb.sweep_scene(noise_amp=PLANET, face=2, root_quad_size=12720000.0, sphere=1)
    e, q = plb.make_requests_range(sc, 5, 40, 8)
    tiles = [(5,) + plb.morton_decode(40 + i) for i in range(8)]
    e1 = plb.elev_make_reqs(tiles, root_quad_size=12720000.0, noise_amp=PLANET, face=2)
    q1 = plb.norm_make_reqs(tiles, sc.norm, root_quad_size=12720000.0)
    for name in ("dx", "dy", "noise_r", "noise_l", "rs", "pixel_size", "level", "tx", "ty"):
        np.testing.assert_array_equal(e[name], e1[name])
    for name in ("deform", "corners", "verticals", "norms", "w2t", "p2t", "smooth"):
        np.testing.assert_array_equal(q[name], q1[name])


def test_noise_layers_match_oracle(plb, oracle):
    """pl_noise_init's host half (createDemNoise + R16F rounding); needs no device up to the upload"""
    if not _has_gpu():
        pytest.skip("pl_noise_init uploads to the device")
    with plb.Context(0) as ctx:
        np.testing.assert_array_equal(ctx.noise_init(101), oracle.dem_noise(101))


def test_sweep_plan_covers_the_planet_once(plb):
    import sweep
    units = sweep.planet_units()
    assert len(units) == 96 and sweep.pairs_in_units(units, 10) == 6 * (4 ** 11 - 1) // 3 == 8388606
    off, cap = sweep.region_offsets(10)
    assert cap == 21 + sum(4 ** d for d in range(1, 9))
    seen = set()
    for f, level, m0, n, s0, p0, pm0 in sweep.batches(units, 6):
        assert s0 + n <= sweep.region_offsets(6)[1]
        for i in range(n):
            key = (f, level) + sweep.morton_decode(m0 + i)
            if level > 2:
                assert key not in seen
            seen.add(key)
            if level > 0:       # the parent was produced before and sits where the plan says
                tx, ty = key[2], key[3]
                assert (f, level - 1, tx // 2, ty // 2) in seen
    assert len(seen) == 6 * (4 ** 7 - 1) // 3
    for world in (1, 2, 4, 8):
        parts = [sweep.units_of_rank(units, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == sorted(units) and len({len(p) for p in parts}) == 1
    assert [plb.morton_decode(plb.morton_encode(x, y)) for x, y in ((0, 0), (5, 9), (1023, 77))] == [(0, 0), (5, 9), (1023, 77)]


def test_residual_file_writer_matches_the_reference_format(plb, tmp_path):
    """pl_residual_write_file (HeightMipmap::generate / produceTile): the oracle's reader -- pinned on the
    reference's own DEM.dat -- reads every tile back; all-zero tiles share one blob; the blobs of a level
    lie in Lebesgue order; the header and the offset-table size are the reference's."""
    import struct
    import numpy as np
    import orc
    import resid_synth as rs
    min_level, max_level, tile_size = 3, 5, 192
    _, tiles = rs.container(min_level=min_level, max_level=max_level, tile_size=tile_size, zero_fraction=0.3, seed=5)
    path = str(tmp_path / "DEM.dat")
    plb.residual_write_file(path, tiles, min_level, max_level, tile_size, root=(0, 0, 0), scale=0.5)
    data = open(path, "rb").read()
    head = struct.unpack("<6if", data[:28])
    assert head[:6] == (min_level, max_level, tile_size, 0, 0, 0) and head[6] == 0.5
    nt = rs.n_tiles(min_level, max_level)
    offs = np.frombuffer(data[28:28 + 8 * nt], "<u4").reshape(nt, 2)
    assert 28 + 8 * nt + int(offs[:, 1].max()) == len(data)
    rd = orc.Resid(data)
    for tid, t in tiles.items():
        raw, w, h = rd.inflate(tid)
        assert (w, h) == t.shape and np.array_equal(np.frombuffer(raw, "<i2").reshape(t.shape), t), tid
    zero_ids = [tid for tid, t in tiles.items() if not t.any()]
    assert len(zero_ids) > 2 and len({tuple(offs[t]) for t in zero_ids}) == 1
    # Lebesgue order inside a level: offsets grow along the Z curve (skipping the shared zero blob)
    l = max_level - min_level
    seen = []
    for m in range(4 ** l):
        tx = sum(((m >> (2 * b)) & 1) << b for b in range(l))
        ty = sum(((m >> (2 * b + 1)) & 1) << b for b in range(l))
        tid = rs.tile_id(min_level, max_level, tx, ty)
        if tiles[tid].any():
            seen.append(int(offs[tid, 0]))
    assert seen == sorted(seen) and len(set(seen)) == len(seen)
    with pytest.raises(plb.PlError) as e:
        plb.residual_write_file(str(tmp_path / "no" / "such" / "dir.dat"), tiles, min_level, max_level, tile_size)
    assert e.value.code == plb.PL_ERR_IO
