"""GPU parity tests THROUGH the C++ host layer (namespace proland: ResourceManager -> TileCache ->
Elevation/Normal/ResidualProducer -> BatchScheduler -> C ABI -> CUDA kernels), the way a Proland
application drives the path: XML archive, getTile, Scheduler::run, read the slots.  Every tile is
compared bit for bit with the CPU oracle."""
import os
import struct

import numpy as np
import pytest

import quadtree as qt
import resid_synth as rs

pytestmark = pytest.mark.gpu

FRACTAL = [-140, -100, -15, -8, 5, 2.5, 1.5, 1, 0.5, 0.25, 0.1, 0.05]
PLANET = [-3250, -1590, -1125, -795, -561, -397, -140, -100, 15, 8, 5, 2.5, 1.5, 1, 0.5, 0.25, 0.1, 0.05]


@pytest.fixture(scope="module")
def ph():
    import proland_host
    if not os.path.exists(proland_host.LIB_PATH):
        proland_host.build()
    return proland_host


def terrain_archive(n_tiles=512, noise=FRACTAL, elev_filter="LINEAR", normal_format="RG8", extra_elev="", extra_norm="",
                    name="groundElevations1", prefetch=""):
    return """<?xml version="1.0" ?>
<archive>
    <multithreadScheduler name="defaultScheduler" nthreads="3" fps="0" %s/>
    <tileCache name="groundElevations" scheduler="defaultScheduler">
        <gpuTileStorage tileSize="101" nTiles="%d"
            internalformat="RGB32F" format="RGB" type="FLOAT" min="%s" mag="%s"/>
    </tileCache>
    <elevationProducer name="%s" cache="groundElevations" noise="%s" %s/>
    <tileCache name="groundNormals" scheduler="defaultScheduler">
        <gpuTileStorage tileSize="97" nTiles="%d"
            internalformat="%s" format="RG" type="FLOAT" min="LINEAR" mag="LINEAR"/>
    </tileCache>
    <normalProducer name="groundNormals1" cache="groundNormals" elevations="%s" %s/>
    <terrainNode name="terrain" size="50000" zmin="0" zmax="5000" splitFactor="2" maxLevel="16"/>
</archive>""" % (prefetch, n_tiles, elev_filter, elev_filter, name, ",".join(str(a) for a in noise), extra_elev, n_tiles,
                 normal_format, name, extra_norm)


def check_against(ref, normals, elevations, tiles):
    for t in tiles:
        e, n, mm = ref[(t.level, t.tx, t.ty)]
        assert np.array_equal(t.download(), n), (t.level, t.tx, t.ty)
        et = elevations.find_tile(t.level, t.tx, t.ty, include_cache=True, done=True)
        assert et is not None
        assert np.array_equal(et.download(), e), (t.level, t.tx, t.ty)
        assert et.minmax() == mm


def test_fractalterrain_archive_levels_0_3(ph, oracle):
    """config 1 through the plugin surface; one kernel launch per producer per quadtree level.
    As in the reference's fractalterrain.xml the producer is called groundElevations1, and a name
    ending in 1..6 selects that cube face's noise lattice (ElevationProducer.cpp:503-509): face 1."""
    ref = qt.oracle_quadtree(oracle, 3, noise_amp=FRACTAL, face=1)
    with ph.Scene(terrain_archive()) as scene:
        normals, elevations = scene.producer("groundNormals1"), scene.producer("groundElevations1")
        assert (elevations.type, elevations.task_type) == ("ElevationProducer", "CreateElevationTile")
        assert (normals.type, normals.task_type) == ("NormalProducer", "CreateNormalTile")
        assert elevations.info()["border"] == 2 and normals.info()["border"] == 0 and normals.info()["referenced"] == 1
        normals.set_root_quad_size(100000.0)          # TerrainNode: 2 * size, passed on to the elevations
        launches0 = ph.lib().plh_device_launches(-1)
        tiles = [normals.get_tile(3, tx, ty) for ty in range(8) for tx in range(8)]
        scene.scheduler("defaultScheduler").run(tiles)
        assert all(t.done for t in tiles)
        # 85 tiles each, 4 batches each (levels 0..3): 8 launches for 170 tiles
        assert elevations.counts() == (85, 4) and normals.counts() == (85, 4)
        assert ph.lib().plh_device_launches(-1) - launches0 == 8
        check_against(ref, normals, elevations, tiles)
        # ancestors are in the cache (unused) and identical too
        for key in [(0, 0, 0), (1, 1, 0), (2, 3, 1)]:
            nt = normals.find_tile(*key, include_cache=True, done=True)
            assert nt is not None and np.array_equal(nt.download(), ref[key][1])
        st = scene.cache("groundNormals").stats()
        assert st["used"] == 64 and st["unused"] == 21
        for t in tiles:
            normals.put_tile(t)
        assert scene.cache("groundNormals").stats()["used"] == 0
        assert scene.cache("groundElevations").stats()["used"] == 0


@pytest.mark.parametrize("face", [1, 4, 6])
def test_fractalplanet_archive_face_from_name_suffix(ph, oracle, face):
    """config 2: the producer name's last digit is the cube face (ElevationProducer.cpp:503-509), sphere normals"""
    kw = dict(noise_amp=PLANET, face=face, root_quad_size=12720000.0, sphere=1)
    ref = qt.oracle_quadtree(oracle, 2, **kw)
    xml = terrain_archive(noise=PLANET, name="groundElevations%d" % face, extra_norm='deform="sphere"')
    with ph.Scene(xml) as scene:
        normals, elevations = scene.producer("groundNormals1"), scene.producer("groundElevations%d" % face)
        normals.set_root_quad_size(12720000.0)
        tiles = [normals.get_tile(2, tx, ty) for ty in range(4) for tx in range(4)]
        scene.scheduler("defaultScheduler").run(tiles)
        check_against(ref, normals, elevations, tiles)
        for t in tiles:
            normals.put_tile(t)


def test_shader_variant_flip_and_gridsize_attributes(ph, oracle):
    ref = qt.oracle_quadtree(oracle, 2, noise_amp=FRACTAL[:2] + [30, 20], noise_mode=0, flip=1, no_clamp=1, elev_filter=0)
    xml = terrain_archive(noise=FRACTAL[:2] + [30, 20], elev_filter="NEAREST",
                          extra_elev='flip="true" upsampleProg="upsampleShader-plain-noClamp;" gridSize="24" face="0"')
    with ph.Scene(xml) as scene:
        normals, elevations = scene.producer("groundNormals1"), scene.producer("groundElevations1")
        normals.set_root_quad_size(100000.0)
        tiles = [normals.get_tile(2, tx, ty) for ty in range(4) for tx in range(4)]
        scene.scheduler("defaultScheduler").run(tiles)
        check_against(ref, normals, elevations, tiles)
        for t in tiles:
            normals.put_tile(t)


def test_small_cache_evicts_recomputes_and_invalidates(ph, oracle):
    """TileCache under pressure: 24 slots for a 85-tile quadtree walked quad by quad"""
    ref = qt.oracle_quadtree(oracle, 3, noise_amp=FRACTAL, face=1)
    with ph.Scene(terrain_archive(n_tiles=24)) as scene:
        normals, elevations = scene.producer("groundNormals1"), scene.producer("groundElevations1")
        normals.set_root_quad_size(100000.0)
        sched = scene.scheduler("defaultScheduler")
        for rep in range(2):
            for qy in range(4):
                for qx in range(4):
                    tiles = [normals.get_tile(3, 2 * qx + i, 2 * qy + j) for j in range(2) for i in range(2)]
                    sched.run(tiles)
                    check_against(ref, normals, elevations, tiles)
                    for t in tiles:
                        normals.put_tile(t)
        st = scene.cache("groundNormals").stats()
        assert st["free"] == 0 and st["used"] == 0 and st["unused"] == 24
        assert st["misses"] > 85                      # tiles were evicted and made again
        assert normals.counts()[0] == st["misses"]
        # cache full: 24 slots cannot hold the 64 + ancestors a whole level needs at once
        ph.lib().plh_quiet_errors(1)
        held = []
        with pytest.raises((ph.HostError, AssertionError)):
            for ty in range(8):
                for tx in range(8):
                    held.append(normals.get_tile(3, tx, ty))
        ph.lib().plh_quiet_errors(0)
        # what was acquired is usable: tasks that hold input tiles must run (or die) before the scene closes
        sched.run(held)
        check_against(ref, normals, elevations, held)
        for t in held:
            normals.put_tile(t)
        # invalidating the elevations re-runs them and the normals that depend on them
        t = normals.get_tile(1, 1, 1)
        sched.run([t])
        before = elevations.counts()[0], normals.counts()[0]
        elevations.invalidate_tiles()
        sched.run([t])
        after = elevations.counts()[0], normals.counts()[0]
        assert after[0] - before[0] >= 2 and after[1] - before[1] >= 1
        assert np.array_equal(t.download(), ref[(1, 1, 1)][1])
        normals.put_tile(t)


def test_prefetch_produces_tiles_ahead(ph, oracle):
    ref = qt.oracle_quadtree(oracle, 2, noise_amp=FRACTAL, face=1)
    with ph.Scene(terrain_archive(prefetch='prefetchRate="2" prefetchQueue="64"')) as scene:
        normals = scene.producer("groundNormals1")
        normals.set_root_quad_size(100000.0)
        sched = scene.scheduler("defaultScheduler")
        assert normals.prefetch_tile(2, 1, 3) and normals.prefetch_tile(2, 2, 0) and normals.prefetch_tile(2, 3, 3)
        sched.run([])                                   # two prefetch tasks per frame
        done = [k for k in [(2, 1, 3), (2, 2, 0), (2, 3, 3)] if normals.find_tile(*k, include_cache=True, done=True)]
        assert len(done) == 2
        sched.run([])
        for k in [(2, 1, 3), (2, 2, 0), (2, 3, 3)]:
            t = normals.find_tile(*k, include_cache=True, done=True)
            assert t is not None and np.array_equal(t.download(), ref[k][1])


def test_rgba8_normal_storage_from_archive(ph, oracle):
    """internalformat="RGBA8": fine + coarse normals (format code 1), parent normal tiles are real inputs"""
    with ph.Scene(terrain_archive(normal_format="RGBA8")) as scene:
        normals, elevations = scene.producer("groundNormals1", channels=4), scene.producer("groundElevations1")
        normals.set_root_quad_size(100000.0)
        tiles = [normals.get_tile(2, tx, ty) for ty in range(4) for tx in range(4)]
        scene.scheduler("defaultScheduler").run(tiles)
        ref = {}
        for level in range(3):
            for (_, tx, ty) in qt.level_tiles(level):
                e = elevations.find_tile(level, tx, ty, include_cache=True, done=True).download()
                p = oracle.normal_uniforms(level, tx, ty, components=4, rootQuadSize=100000.0)
                parent = ref[(level - 1, tx // 2, ty // 2)].astype(np.float32) / np.float32(255.0) if level else None
                ref[(level, tx, ty)] = oracle.pack_unorm8(oracle.normal_tile(p, e, parent), 4)
        for t in tiles:
            assert np.array_equal(t.download(), ref[(2, t.tx, t.ty)])
            normals.put_tile(t)


SRTM = """<?xml version="1.0" ?>
<archive>
    <multithreadScheduler name="defaultScheduler" nthreads="3" fps="0"/>
    <tileCache name="groundResiduals" scheduler="defaultScheduler">
        <cpuFloatTileStorage tileSize="197" channels="1" capacity="64"/>
    </tileCache>
    <residualProducer name="groundResiduals2" cache="groundResiduals" file="dem/DEM2.dat" delta="2" scale="0.5">
        <residualProducer name="groundResiduals2a" cache="groundResiduals" file="dem/DEM2a.dat"/>
    </residualProducer>
    <tileCache name="groundElevations" scheduler="defaultScheduler">
        <gpuTileStorage tileSize="101" nTiles="1296"
            internalformat="RGB32F" format="RGB" type="FLOAT" min="NEAREST" mag="NEAREST"/>
    </tileCache>
    <elevationProducer name="groundElevations2" cache="groundElevations" residuals="groundResiduals2" flip="true"
        noise="0,0,0,5,2.5,1,0.5,0.25,0.1,0.05,0.025,0.01,0.01,0.005,0.005"/>
    <tileCache name="groundNormals" scheduler="defaultScheduler">
        <gpuTileStorage tileSize="97" nTiles="1296"
            internalformat="RG8" format="RG" type="FLOAT" min="LINEAR" mag="LINEAR"/>
    </tileCache>
    <normalProducer name="groundNormals2" cache="groundNormals" elevations="groundElevations2" deform="sphere"/>
</archive>"""


def test_earth_srtm_archive_residual_files_delta_and_nested_producer(ph, oracle, tmp_path):
    """config 3: residual file in the reference's container format (DEM.dat geometry: minLevel 3,
    tileSize 192), delta = 2 root composition, a nested residualProducer for a sub-pyramid, flip,
    NEAREST elevation storage, sphere normals -- ResidualProducer -> ElevationProducer -> NormalProducer"""
    os.makedirs(tmp_path / "dem")
    main, _ = rs.container(min_level=3, max_level=4, tile_size=192, scale=2.0, seed=5)
    # sub-pyramid under stored tile (5, 1, 2) of the main file: stored levels 5..6
    child, _ = rs.container(min_level=0, max_level=1, tile_size=192, root=(5, 1, 2), scale=2.0, seed=6,
                            amps=(3.9, 1.0), zero_fraction=0.0)
    (tmp_path / "dem" / "DEM2.dat").write_bytes(main)
    (tmp_path / "dem" / "DEM2a.dat").write_bytes(child)
    res = oracle.Resid(main, delta=2, zscale=0.5)
    res.add_child(oracle.Resid(child, delta=0, zscale=1.0))
    amp = [0, 0, 0, 5, 2.5, 1, 0.5, 0.25, 0.1, 0.05, 0.025, 0.01, 0.01, 0.005, 0.005]
    scene_o = oracle.make_scene(W=101, rootQuadSize=12720000.0, face=2, flip=1, noiseAmp=amp, sphere=1, elev_filter=0, resid=res)
    noise = oracle.dem_noise(101)
    max_level = 4
    # the quadtree branch below elevation tile (3, 2, 4) .. (3, 3, 5) reaches the nested file
    want = [(4, tx, ty) for ty in range(8, 12) for tx in range(4, 8)] + [(2, tx, ty) for ty in range(4) for tx in range(4)]
    ref = {}

    def make(level, tx, ty):
        if (level, tx, ty) in ref:
            return ref[(level, tx, ty)]
        parent = make(level - 1, tx // 2, ty // 2)[0] if level else None
        rt = res.create_tile(level, tx // 2, ty // 2) if res.has_tile(level, tx // 2, ty // 2) else None
        ref[(level, tx, ty)] = oracle.produce_pair(scene_o, noise, level, tx, ty, parent, rt)
        return ref[(level, tx, ty)]

    with ph.Scene(SRTM, data_dir=str(tmp_path)) as scene:
        normals, elevations = scene.producer("groundNormals2"), scene.producer("groundElevations2")
        residuals = scene.producer("groundResiduals2")
        assert residuals.residual_info() == dict(min_level=3, max_level=4, delta=2)
        assert residuals.info()["border"] == 2 and residuals.type == "ResidualProducer"
        assert residuals.residual_tile_size(0) == 24 and residuals.residual_tile_size(5) == 192
        assert residuals.residual_tile_id(4, 1, 1) == res.tile_id(4, 1, 1)
        # stored levels 0..4 = elevation levels -2..2; level 3 / 4 only under the nested file's root
        assert residuals.has_tile(0, 0, 0) and residuals.has_tile(2, 1, 1) and not residuals.has_tile(3, 0, 0)
        assert residuals.has_tile(3, 1, 2) and residuals.has_tile(4, 3, 5) and not residuals.has_tile(5, 6, 10)
        normals.set_root_quad_size(12720000.0)
        tiles = [normals.get_tile(*k) for k in want]
        scene.scheduler("defaultScheduler").run(tiles)
        assert any(res.has_tile(4, t.tx // 2, t.ty // 2) for t in tiles)
        for t in tiles:
            e, n = make(t.level, t.tx, t.ty)
            assert np.array_equal(t.download(), n), (t.level, t.tx, t.ty)
            et = elevations.find_tile(t.level, t.tx, t.ty, include_cache=True, done=True)
            assert np.array_equal(et.download(), e), (t.level, t.tx, t.ty)
        # the composed root residual tile itself (float, 101 of 197 texels wide)
        rt = residuals.find_tile(0, 0, 0, include_cache=True, done=True)
        w = res.tile_size(2) + 5
        assert np.array_equal(rt.download()[:w, :w], res.create_tile(0, 0, 0)[:w, :w])
        for t in tiles:
            normals.put_tile(t)


def test_missing_residual_file_has_no_tiles(ph):
    """ResidualProducer.cpp:86-94: fopen fails -> error logged, maxLevel = -1, no tile anywhere"""
    ph.lib().plh_quiet_errors(1)
    with ph.Scene(SRTM, data_dir="/nonexistent") as scene:
        r = scene.producer("groundResiduals2")
        assert r.residual_info()["max_level"] == -1 and not r.has_tile(0, 0, 0)
    ph.lib().plh_quiet_errors(0)


def test_debug_log_counts_tiles_like_the_reference(ph):
    """the reference's only production counter: DEBUG log lines "Elevation tile ..." / "Normal tile ..." """
    ph.lib().plh_debug_log(1, 0)
    n0 = ph.lib().plh_debug_log_lines()
    with ph.Scene(terrain_archive()) as scene:
        normals = scene.producer("groundNormals1")
        normals.set_root_quad_size(100000.0)
        t = normals.get_tile(1, 0, 1)
        scene.scheduler("defaultScheduler").run([t])
        normals.put_tile(t)
    lines = ph.lib().plh_debug_log_lines() - n0
    ph.lib().plh_debug_log(0, 0)
    assert lines >= 4          # 2 elevation + 2 normal tiles (+ cache statistics lines)
