"""GPU parity of the ortho path (OrthoProducer / upsampleOrthoShader): pl_ortho_batch
through the C ABI against the CPU oracle (oracle/orc_ortho.c), byte for byte."""
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ORTHO = json.load(open(os.path.join(HERE, "golden", "ortho.json")))
TERRAIN3 = dict(hsv=1, cnoise=(70, 80, 100), rnoise=(60, 150, 20), noise_amp=[255] * 17, face=1)


def _sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def _gpu_quadtree(plb, ctx, sc, max_level):
    """levels 0..max_level, one batch per level, slots in level order / Morton order inside a level"""
    W = sc.tile_w
    n = (4 ** (max_level + 1) - 1) // 3
    pool = ctx.pool(plb.POOL_ORTHO, W, n)
    ctx.ortho_noise_init(W)
    off = [(4 ** l - 1) // 3 for l in range(max_level + 2)]
    for l in range(max_level + 1):
        reqs = plb.ortho_make_requests_range(sc, l, 0, 4 ** l, out_slot0=off[l], parent_slot0=off[l - 1] if l else 0)
        ctx.ortho_batch(sc, pool, None, reqs)
    ctx.sync()
    return np.stack([pool.download(s) for s in range(n)])


def _oracle_quadtree(oracle, sc, max_level):
    return oracle.ortho_quadtree(max_level, W=sc.tile_w, face=sc.face, noise_amp=list(sc.noise_amp)[:sc.n_amp],
                                 noise_color=list(sc.noise_color), root_noise_color=list(sc.root_noise_color),
                                 hsv=sc.hsv, scale=sc.scale)


def _assert_same(gpu, ref, what):
    bad = np.argwhere(gpu != ref)
    assert bad.size == 0, "%s: %d bytes differ from the oracle, first at %s: gpu %d, oracle %d" % (
        what, len(bad), tuple(bad[0]), gpu[tuple(bad[0])], ref[tuple(bad[0])])


def test_ortho_noise_on_device_matches_oracle(plb, ctx, oracle):
    for W in (196, 100):
        assert np.array_equal(ctx.ortho_noise_init(W, want_host=True), oracle.ortho_noise(W))


def test_terrain3_hsv_quadtree(plb, ctx, oracle):
    """terrain3/helloworld.xml:47-49: hsv noise, rnoise 60,150,20, cnoise 70,80,100, amplitudes 255: levels 0..4"""
    sc = plb.ortho_scene(**TERRAIN3)
    gpu = _gpu_quadtree(plb, ctx, sc, 4)
    _assert_same(gpu, _oracle_quadtree(oracle, sc, 4), "terrain3 hsv")
    assert [_sha(t) for t in gpu[:85]] == ORTHO["terrain3_hsv"]["levels_0_3_sha1"]


@pytest.mark.parametrize("face", [1, 2, 5, 6])
def test_plain_noise_quadtree_per_face(plb, ctx, oracle, face):
    """hsv="false": noiseColor * scale * amplitude added per channel; the layer / rotation choice per cube face"""
    sc = plb.ortho_scene(hsv=0, cnoise=(127.5, 40, 90, 10), noise_amp=[0, 255, 255, 128, 64], face=face, scale=2.0)
    gpu = _gpu_quadtree(plb, ctx, sc, 3)
    _assert_same(gpu, _oracle_quadtree(oracle, sc, 3), "plain noise, face %d" % face)


def test_plain_golden(plb, ctx):
    sc = plb.ortho_scene(hsv=0, cnoise=(127.5, 0, 0, 0), noise_amp=[0, 255, 255, 255, 255], face=3)
    gpu = _gpu_quadtree(plb, ctx, sc, 3)
    assert [_sha(t) for t in gpu] == ORTHO["plain"]["levels_0_3_sha1"]


@pytest.mark.parametrize("hsv", [0, 1])
def test_rgb8_storage_skips_alpha(plb, ctx, oracle, hsv):
    """out_channels = 3 (an RGB8 storage, terrain3/helloworld.xml:43): the colour channels are the oracle's, the
    alpha byte is not computed and reads 0"""
    kw = dict(hsv=hsv, cnoise=(70, 80, 100, 90), rnoise=(60, 150, 20, 200), noise_amp=[255] * 6, face=4)
    gpu = _gpu_quadtree(plb, ctx, plb.ortho_scene(out_channels=3, **kw), 3)
    ref = _oracle_quadtree(oracle, plb.ortho_scene(**kw), 3)
    _assert_same(gpu[..., :3], ref[..., :3], "rgb8 hsv=%d" % hsv)
    assert (gpu[..., 3] == 0).all()


def test_tile_w_100(plb, ctx, oracle):
    """the 100-texel storages of the land-cover producers (exercise2/helloworld.xml:43-52)"""
    sc = plb.ortho_scene(tile_w=100, hsv=1, cnoise=(30, 200, 150, 90), noise_amp=[255, 200, 150, 100], face=2)
    _assert_same(_gpu_quadtree(plb, ctx, sc, 3), _oracle_quadtree(oracle, sc, 3), "tile_w 100")


@pytest.mark.parametrize("hsv,channels", [(0, 4), (0, 3), (1, 3), (1, 4), (0, 1)])
def test_residual_tiles(plb, ctx, oracle, hsv, channels):
    """byte residuals (OrthoCPUProducer tiles uploaded like OrthoProducer.cpp:296-318): random residuals with
    1, 3 and 4 channels, on level 0 (no parent) and on children of a random parent, hsv and plain noise;
    tiles without a residual in the same batch"""
    W = 196
    rng = np.random.default_rng(100 * hsv + channels)
    sc = plb.ortho_scene(tile_w=W, channels=channels, hsv=hsv, cnoise=(70, 80, 100, 60), rnoise=(60, 150, 20, 99),
                         noise_amp=[200, 255, 255], face=1, scale=2.0)
    pool = ctx.pool(plb.POOL_ORTHO, W, 8)
    rpool = ctx.pool(plb.POOL_ORTHO, W, 8)
    nz = ctx.ortho_noise_init(W, want_host=True)
    parent = rng.integers(0, 256, (W, W, 4), dtype=np.uint8)
    parent[:50] = rng.integers(0, 256, 4, dtype=np.uint8)            # a flat area: delta == 0 pixels in the hsv branch
    parent[50:60, :, :3] = 0                                         # black: maxVal == 0
    pool.upload(0, parent)
    tiles = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (1, 0, 1), (1, 1, 1), (1, 1, 1)]
    has = [1, 1, 1, 0, 1, 1]
    resid = []
    for i, h in enumerate(has):
        r = rng.integers(0, 256, (W, W, channels), dtype=np.uint8)
        if i == 4:
            r[:] = 128                                               # the neutral residual
        if i == 5:
            r = (128 + rng.integers(-6, 7, (W, W, channels))).astype(np.uint8)   # what real residual files hold
        resid.append(r if h else None)
        if h:
            full = np.zeros((W, W, 4), np.uint8)
            full[..., :channels] = r
            full[..., channels:] = 37                                # garbage in the unused channels: must be ignored
            rpool.upload(i, full)
    reqs = plb.ortho_make_reqs(sc, tiles, has)
    for i, q in enumerate(reqs):
        q["out_slot"] = 1 + i
        q["parent_slot"] = 0 if tiles[i][0] > 0 else -1
        q["resid_slot"] = i if has[i] else -1
    ctx.ortho_batch(sc, pool, rpool, reqs)
    ctx.sync()
    for i, (l, tx, ty) in enumerate(tiles):
        p = oracle.ortho_uniforms(l, tx, ty, W=W, face=1, noise_amp=[200, 255, 255], noise_color=list(sc.noise_color),
                                  root_noise_color=list(sc.root_noise_color), hsv=hsv, scale=2.0, has_residual=has[i])
        want = oracle.ortho_tile(p, parent if l > 0 else None, resid[i], nz, channels=channels)
        _assert_same(pool.download(1 + i), want, "tile %d %s hsv=%d channels=%d" % (i, tiles[i], hsv, channels))


def test_ortho_errors(plb, ctx):
    sc = plb.ortho_scene(**TERRAIN3)
    with pytest.raises(plb.PlError) as e:
        ctx.pool(plb.POOL_ORTHO, 197, 4)
    assert e.value.code == plb.PL_ERR_ARG
    pool = ctx.pool(plb.POOL_ORTHO, 196, 4)
    reqs = plb.ortho_make_reqs(sc, [(0, 0, 0)])
    reqs[0]["out_slot"] = 0
    with pytest.raises(plb.PlError):                                  # noise not initialised
        ctx.ortho_batch(sc, pool, None, reqs)
    ctx.ortho_noise_init(196)
    reqs[0]["out_slot"] = 4
    with pytest.raises(plb.PlError):                                  # slot out of range
        ctx.ortho_batch(sc, pool, None, reqs)
    reqs[0]["out_slot"] = 0
    reqs[0]["resid_slot"] = 1
    with pytest.raises(plb.PlError):                                  # residual slot without a residual pool
        ctx.ortho_batch(sc, pool, None, reqs)
    reqs[0]["resid_slot"] = -1
    ctx.ortho_batch(sc, pool, None, reqs)
    ctx.ortho_batch(sc, pool, None, reqs[:0])                         # empty batch
    ctx.sync()


def test_ortho_idempotent_and_order_independent(plb, ctx):
    """producing a level twice, or in two halves in the other order, gives the same bytes (tiles are
    independent given their parents)"""
    sc = plb.ortho_scene(**TERRAIN3)
    a = _gpu_quadtree(plb, ctx, sc, 3)
    W, n = 196, 85
    pool = ctx.pool(plb.POOL_ORTHO, W, n)
    off = [0, 1, 5, 21]
    for l in range(4):
        reqs = plb.ortho_make_requests_range(sc, l, 0, 4 ** l, out_slot0=off[l], parent_slot0=off[l - 1] if l else 0)
        h = len(reqs) // 2
        ctx.ortho_batch(sc, pool, None, reqs[h:])
        ctx.ortho_batch(sc, pool, None, reqs[:h])
        ctx.ortho_batch(sc, pool, None, reqs)
    ctx.sync()
    b = np.stack([pool.download(s) for s in range(n)])
    assert np.array_equal(a, b)


# ------------------------------------------------- through the C++ host layer (proland::OrthoProducer)

@pytest.fixture(scope="module")
def ph():
    import proland_host
    if not os.path.exists(proland_host.LIB_PATH):
        proland_host.build()
    return proland_host


TERRAIN3_XML = """<?xml version="1.0" ?>
<archive>
    <multithreadScheduler name="defaultScheduler" nthreads="3" fps="0"/>
    <tileCache name="groundOrthoGpu" scheduler="defaultScheduler">
        <gpuTileStorage tileSize="196" nTiles="512"
            internalformat="%s" format="RGB" type="UNSIGNED_BYTE" min="LINEAR_MIPMAP_LINEAR" mag="LINEAR"
            anisotropy="16"/>
    </tileCache>
    <orthoProducer name="groundOrthoGpu%d" cache="groundOrthoGpu"
        hsv="true" rnoise="60,150,20" cnoise="70,80,100"
        noise="255,255,255,255,255,255,255,255,255,255,255,255,255,255,255,255,255"/>
</archive>"""


@pytest.mark.parametrize("fmt,face", [("RGB8", 1), ("RGBA8", 5)])
def test_terrain3_archive_through_the_host_layer(ph, plb, oracle, fmt, face):
    """terrain3/helloworld.xml:42-49 verbatim (storage + orthoProducer): XML -> TileCache -> OrthoProducer ->
    BatchScheduler -> pl_ortho_batch; one launch per quadtree level; every tile equals the oracle's.
    rnoise / cnoise with three items: the fourth is atof("") / 255 = 0 (OrthoProducer.cpp:462-481)."""
    kw = dict(W=196, face=face, noise_amp=[255] * 17, hsv=1, scale=2.0,
              noise_color=[np.float32(v) / np.float32(255) for v in (70, 80, 100, 0)],
              root_noise_color=[np.float32(v) / np.float32(255) for v in (60, 150, 20, 0)])
    ref = oracle.ortho_quadtree(3, **kw)
    with ph.Scene(TERRAIN3_XML % (fmt, face)) as scene:
        ortho = scene.producer("groundOrthoGpu%d" % face)
        assert (ortho.type, ortho.task_type) == ("OrthoProducer", "CreateOrthoTile")
        assert ortho.info()["border"] == 2 and ortho.info()["gpu"] and ortho.info()["tile_size"] == 196
        assert ortho.has_tile(20, 0, 0)
        launches0 = ph.lib().plh_device_launches(-1)
        tiles = [ortho.get_tile(3, tx, ty) for ty in range(8) for tx in range(8)]
        scene.scheduler("defaultScheduler").run(tiles)
        assert all(t.done for t in tiles)
        assert ortho.counts() == (85, 4)
        assert ph.lib().plh_device_launches(-1) - launches0 == 4
        nch = 3 if fmt == "RGB8" else 4
        for t in tiles:
            want = ref[21 + plb.morton_encode(t.tx, t.ty)]
            assert np.array_equal(t.download()[..., :nch], want[..., :nch]), (t.tx, t.ty)
        root = ortho.find_tile(0, 0, 0, include_cache=True, done=True)
        assert root is not None and np.array_equal(root.download()[..., :nch], ref[0][..., :nch])
        for t in tiles:
            ortho.put_tile(t)
        assert scene.cache("groundOrthoGpu").stats()["used"] == 0


def test_ortho_producer_max_level(ph):
    xml = TERRAIN3_XML % ("RGBA8", 1)
    xml = xml.replace('hsv="true"', 'hsv="true" maxLevel="7"')
    with ph.Scene(xml) as scene:
        ortho = scene.producer("groundOrthoGpu1")
        assert ortho.has_tile(7, 3, 3) and not ortho.has_tile(8, 3, 3)


# ------------------------------------------------------------- OrthoCPUProducer: residual files on the device

def _ortho_file(channels, max_level=1, seed=3, W=196):
    import resid_synth as rs
    rng = np.random.default_rng(seed)
    tiles = {}
    for l in range(max_level + 1):
        for ty in range(1 << l):
            for tx in range(1 << l):
                t = (128 + rng.integers(-9, 10, (W, W, channels))).astype(np.uint8)
                t[:40] = 128
                tiles[(l, tx, ty)] = t
    return tiles, rs.ortho_container(tiles, max_level, W - 4, channels)


@pytest.fixture(params=[1, 2], ids=["warp-per-stream", "tokenizer+resolver"])
def inflate_path(request, ctx):
    ctx.inflate_path(request.param)
    yield request.param
    ctx.inflate_path(0)


@pytest.mark.parametrize("channels", [1, 2, 3, 4])
def test_ortho_decode_batch_matches_oracle_reader(plb, ctx, oracle, channels, inflate_path):
    """pl_ortho_decode_batch (OrthoCPUProducer.cpp:205-232 on the device) against the oracle's reader of the same
    file, byte for byte, for 1..4 channel files; unused channels of the RGBA8 slot are 0"""
    import resid_synth as rs
    tiles, data = _ortho_file(channels)
    keys = sorted(tiles)
    pool = ctx.pool(plb.POOL_ORTHO, 196, len(keys))
    got_ch = ctx.ortho_decode(pool, [rs.ortho_container_blob(data, *k) for k in keys], list(range(len(keys))))
    assert got_ch == channels
    for slot, k in enumerate(keys):
        got = pool.download(slot)
        want = oracle.ortho_cpu_read(data, *k)
        assert np.array_equal(got[..., :channels], want), k
        assert (got[..., channels:] == 0).all()


def test_ortho_decode_errors(plb, ctx):
    import resid_synth as rs
    tiles, data = _ortho_file(3)
    pool = ctx.pool(plb.POOL_ORTHO, 196, 2)
    good = rs.ortho_container_blob(data, 0, 0, 0)
    bad = bytearray(good)
    bad[40:60] = b"\xff" * 20                                       # garbage in the DEFLATE stream
    with pytest.raises(plb.PlError) as e:
        ctx.ortho_decode(pool, [bytes(bad)], [0])
    assert e.value.code == plb.PL_ERR_CORRUPT
    with pytest.raises(plb.PlError) as e:                           # a 100-texel tile into a 196-texel pool
        ctx.ortho_decode(pool, [rs.ortho_tiff_blob(np.zeros((100, 100, 3), np.uint8))], [0])
    assert e.value.code == plb.PL_ERR_CORRUPT
    # an elevation residual blob is 2 x 8-bit samples per texel (ResidualProducer's int16): a legal 2-channel byte tile
    assert ctx.ortho_decode(pool, [rs.tiff_blob(np.full((196, 196), 0x0201, np.int16))], [0]) == 2
    assert (pool.download(0) == np.array([1, 2, 0, 0], np.uint8)).all()
    with pytest.raises(plb.PlError):
        ctx.ortho_decode(pool, [good], [2])                         # slot out of range
    rpool = ctx.pool(plb.POOL_RESID_I16, 197, 2)
    with pytest.raises(plb.PlError):                                # the residual entry point refuses byte pools and v.v.
        ctx.residual_decode(pool, [good], [196], [0])
    with pytest.raises(plb.PlError):
        ctx.ortho_decode(rpool, [good], [0])
    assert ctx.ortho_decode(pool, [good], [1]) == 3


@pytest.mark.parametrize("hsv", [0, 1])
def test_ortho_file_to_tiles(plb, ctx, oracle, hsv):
    """the whole ortho chain of earth-like archives: residual file -> device decode -> OrthoProducer pass with
    residuals on every tile of levels 0..2, against the oracle fed by its own reader"""
    import resid_synth as rs
    W, L = 196, 2
    tiles, data = _ortho_file(3, max_level=L, seed=11)
    keys = [(l, tx, ty) for l in range(L + 1) for ty in range(1 << l) for tx in range(1 << l)]
    n = len(keys)
    sc = plb.ortho_scene(tile_w=W, channels=3, hsv=hsv, cnoise=(70, 80, 100, 60), rnoise=(60, 150, 20, 99),
                         noise_amp=[0, 30, 60], face=2, scale=2.0)
    pool, rpool = ctx.pool(plb.POOL_ORTHO, W, n), ctx.pool(plb.POOL_ORTHO, W, n)
    nz = ctx.ortho_noise_init(W, want_host=True)
    assert ctx.ortho_decode(rpool, [rs.ortho_container_blob(data, *k) for k in keys], list(range(n))) == 3
    slot = {k: i for i, k in enumerate(keys)}
    for l in range(L + 1):
        lk = [k for k in keys if k[0] == l]
        reqs = plb.ortho_make_reqs(sc, lk, [1] * len(lk))
        for q, k in zip(reqs, lk):
            q["out_slot"] = slot[k]
            q["resid_slot"] = slot[k]
            q["parent_slot"] = slot[(l - 1, k[1] // 2, k[2] // 2)] if l else -1
        ctx.ortho_batch(sc, pool, rpool, reqs)
    ctx.sync()
    ref = {}
    for k in keys:
        l, tx, ty = k
        p = oracle.ortho_uniforms(l, tx, ty, W=W, face=2, noise_amp=[0, 30, 60], noise_color=list(sc.noise_color),
                                  root_noise_color=list(sc.root_noise_color), hsv=hsv, scale=2.0, has_residual=1)
        parent = ref[(l - 1, tx // 2, ty // 2)] if l else None
        ref[k] = oracle.ortho_tile(p, parent, oracle.ortho_cpu_read(data, *k), nz, channels=3)
        _assert_same(pool.download(slot[k]), ref[k], "tile %s" % (k,))


ORTHO_FILE_XML = """<?xml version="1.0" ?>
<archive>
    <multithreadScheduler name="defaultScheduler" nthreads="3" fps="0"/>
    <tileCache name="groundOrthoCpu" scheduler="defaultScheduler">
        <cpuByteTileStorage tileSize="196" channels="%d" capacity="64"/>
    </tileCache>
    <orthoCpuProducer name="groundOrthoCpu2" cache="groundOrthoCpu" file="RGB2.dat"/>
    <tileCache name="groundOrthoGpu" scheduler="defaultScheduler">
        <gpuTileStorage tileSize="196" nTiles="128"
            internalformat="RGBA8" format="RGBA" type="UNSIGNED_BYTE" min="LINEAR_MIPMAP_LINEAR" mag="LINEAR"
            anisotropy="16"/>
    </tileCache>
    <orthoProducer name="groundOrthoGpu2" cache="groundOrthoGpu" residuals="groundOrthoCpu2"
        cnoise="70,80,100,60" rnoise="60,150,20,99" noise="0,30,60,90" scale="2" hsv="%s"/>
</archive>"""


@pytest.mark.parametrize("channels,hsv", [(3, "true"), (4, "false")])
def test_ortho_file_archive_through_the_host_layer(ph, plb, oracle, tmp_path, channels, hsv):
    """an earth-style ortho archive (preprocess/helloworld.xml:61-64 + an orthoProducer with residuals): XML ->
    cpuByteTileStorage + orthoCpuProducer (file mapped, blobs inflated on the device) -> orthoProducer.  The file
    has levels 0..2; level-3 tiles have no residual (OrthoCPUProducer::hasTile) and are pure upsample + noise."""
    W, L = 196, 2
    tiles, data = _ortho_file(channels, max_level=L, seed=5)
    (tmp_path / "RGB2.dat").write_bytes(data)
    amp = [0, 30, 60, 90]
    nz = oracle.ortho_noise(W)
    ncol = [np.float32(v) / np.float32(255) for v in (70, 80, 100, 60)]
    rcol = [np.float32(v) / np.float32(255) for v in (60, 150, 20, 99)]
    ref = {}
    chain = [(0, 0, 0), (1, 1, 0), (2, 3, 1), (3, 6, 2)]
    for (l, tx, ty) in chain:
        has = l <= L
        p = oracle.ortho_uniforms(l, tx, ty, W=W, face=2, noise_amp=amp, noise_color=ncol, root_noise_color=rcol,
                                  hsv=int(hsv == "true"), scale=2.0, has_residual=int(has))
        parent = ref[(l - 1, tx // 2, ty // 2)] if l else None
        res = oracle.ortho_cpu_read(data, l, tx, ty) if has else None
        ref[(l, tx, ty)] = oracle.ortho_tile(p, parent, res, nz, channels=channels)
    with ph.Scene(ORTHO_FILE_XML % (channels, hsv), data_dir=str(tmp_path)) as scene:
        ortho, cpu = scene.producer("groundOrthoGpu2"), scene.producer("groundOrthoCpu2")
        assert (cpu.type, cpu.task_type) == ("OrthoCPUProducer", "CreateOrthoCPUTile")
        assert cpu.info()["border"] == 2 and ortho.info()["referenced"] == 1
        assert cpu.has_tile(2, 0, 0) and not cpu.has_tile(3, 0, 0) and ortho.has_tile(3, 0, 0)
        t = ortho.get_tile(3, 6, 2)
        scene.scheduler("defaultScheduler").run([t])
        assert t.done and ortho.counts()[0] == 4 and cpu.counts()[0] == 3
        assert np.array_equal(t.download(), ref[(3, 6, 2)])
        for key in chain[:-1]:
            a = ortho.find_tile(*key, include_cache=True, done=True)
            assert a is not None and np.array_equal(a.download(), ref[key]), key
        r = cpu.find_tile(2, 3, 1, include_cache=True, done=True)
        assert r is not None and np.array_equal(r.download()[..., :channels], tiles[(2, 3, 1)])
        ortho.put_tile(t)
        assert scene.cache("groundOrthoGpu").stats()["used"] == 0 and scene.cache("groundOrthoCpu").stats()["used"] == 0


def test_ortho_cpu_producer_missing_file_and_dxt(ph, tmp_path):
    """a missing file: error log + a producer without tiles (maxLevel = -1, OrthoCPUProducer.cpp:80-85);
    a DXT file is refused at load time"""
    ph.lib().plh_quiet_errors(1)
    try:
        with ph.Scene(ORTHO_FILE_XML % (3, "false"), data_dir=str(tmp_path)) as scene:
            assert not scene.producer("groundOrthoCpu2").has_tile(0, 0, 0)
        _, data = _ortho_file(3, max_level=0)
        dxt = bytearray(data)
        dxt[24] = 1
        (tmp_path / "RGB2.dat").write_bytes(bytes(dxt))
        scene = ph.Scene(ORTHO_FILE_XML % (3, "false"), data_dir=str(tmp_path))
        with pytest.raises(ph.HostError):                           # resources load on first use
            scene.producer("groundOrthoCpu2")
        scene.close()
    finally:
        ph.lib().plh_quiet_errors(0)


def test_ortho_seams_at_scale(plb, ctx):
    """size-independent property at a size the oracle does not reach: all 1 024 tiles of level 5 (and their 341
    ancestors) of terrain3's scene -- every pair of neighbouring tiles agrees, byte for byte, on the 4 texels
    they share (the 2-texel border convention, src/terrain/doc/overview.txt:81-88)"""
    sc = plb.ortho_scene(**TERRAIN3)
    L, W = 5, 196
    n = (4 ** (L + 1) - 1) // 3
    pool = ctx.pool(plb.POOL_ORTHO, W, n)
    ctx.ortho_noise_init(W)
    off = [(4 ** l - 1) // 3 for l in range(L + 2)]
    for l in range(L + 1):
        ctx.ortho_batch(sc, pool, None, plb.ortho_make_requests_range(sc, l, 0, 4 ** l, out_slot0=off[l],
                                                                      parent_slot0=off[l - 1] if l else 0))
    ctx.sync()
    side = 1 << L
    row = None
    seams = 0
    for ty in range(side):
        cur = [pool.download(off[L] + plb.morton_encode(tx, ty)) for tx in range(side)]
        for tx in range(side - 1):
            assert np.array_equal(cur[tx][:, W - 4:], cur[tx + 1][:, :4]), (tx, ty)
            seams += 1
        if row is not None:
            for tx in range(side):
                assert np.array_equal(row[tx][W - 4:, :], cur[tx][:4, :]), (tx, ty)
                seams += 1
        row = cur
    assert seams == 2 * side * (side - 1)


def test_ortho_produce_range_device_requests(plb, ctx, oracle):
    """pl_ortho_produce_range: the requests of a Morton range generated on the device are the host's, byte for byte,
    and the tiles are the oracle's"""
    sc = plb.ortho_scene(**dict(TERRAIN3, face=6))
    L, W = 4, 196
    off = [(4 ** l - 1) // 3 for l in range(L + 2)]
    pool = ctx.pool(plb.POOL_ORTHO, W, off[L + 1])
    ctx.ortho_noise_init(W)
    launches0 = ctx.launches
    for l in range(L + 1):
        ctx.ortho_produce_range(sc, pool, l, 0, 4 ** l, off[l], off[l - 1] if l else 0, 0)
    ctx.sync()
    assert ctx.launches - launches0 == 2 * (L + 1)              # request generation + the ortho kernel per level
    dev, _ = ctx.last_requests(4 ** L)
    host = plb.ortho_make_requests_range(sc, L, 0, 4 ** L, out_slot0=off[L], parent_slot0=off[L - 1])
    assert dev.tobytes() == host.tobytes()
    ref = _oracle_quadtree(oracle, sc, L)
    _assert_same(np.stack([pool.download(s) for s in range(off[L + 1])]), ref, "device-generated requests")
    # a partial range in the middle of a level, parents elsewhere
    pool2 = ctx.pool(plb.POOL_ORTHO, W, 200)
    for s in range(16):
        pool2.upload(100 + s, ref[off[2] + s])
    ctx.ortho_produce_range(sc, pool2, 3, 20, 24, 7, 100 + 5, 5)
    ctx.sync()
    for i in range(24):
        assert np.array_equal(pool2.download(7 + i), ref[off[3] + 20 + i]), i
    with pytest.raises(plb.PlError):
        ctx.ortho_produce_range(sc, pool2, 3, 60, 8, 0, 100, 0)         # Morton range exceeds the level
    with pytest.raises(plb.PlError):
        ctx.ortho_produce_range(sc, pool2, 3, 0, 8, 196, 100, 0)        # slots exceed the pool
