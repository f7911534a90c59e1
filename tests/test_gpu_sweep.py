"""Config 4 (full-subtree batch sweep at level 14) on the GPU: sampled leaf tiles against the oracle's
15-level parent chains, and invariance of the result under the subtree partition (1 / 2 / 4 ranks)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import resid_synth as rs

pytestmark = pytest.mark.gpu

FRACTAL = [-140, -100, -15, -8, 5, 2.5, 1.5, 1, 0.5, 0.25, 0.1, 0.05]


def oracle_chain(oracle, scene, noise, level, tx, ty, resid_of=None):
    parent = None
    for l in range(level + 1):
        x, y = tx >> (level - l), ty >> (level - l)
        rt = resid_of(l, x, y) if resid_of else None
        e, n = oracle.produce_pair(scene, noise, l, x, y, parent, rt)
        parent = e
    return e, n


def test_subtree_sweep_d6_leaves_match_the_oracle(plb, ctx, oracle):
    import subtree_sweep as ss
    import sweep
    d = 6
    plan = sweep.SubtreeSweep(14 - d, (1 << 8) // 3, (1 << 8) // 5, d, unit_depth=4)
    first = plan.root_morton << (2 * d)
    want = [first, first + 1, first + 4 ** d // 2 + 77, first + 4 ** d - 1]
    keep = {"want": want}
    n, fp, plan = ss.run_sweep(plb, ctx, d, unit_depth=4, keep=keep)
    assert n == plan.total_tiles() == 8 + sum(4 ** j for j in range(d + 1))
    scene = oracle.make_scene(W=101, rootQuadSize=100000.0, face=0, noiseAmp=FRACTAL, sphere=0, elev_filter=1)
    noise = oracle.dem_noise(101)
    for m in want:
        tx, ty = plb.morton_decode(m)
        e, nrm = oracle_chain(oracle, scene, noise, 14, tx, ty)
        assert np.array_equal(keep[m][0], e), (tx, ty)
        assert np.array_equal(keep[m][1], nrm), (tx, ty)
    assert fp["lo"] >= 0.0 and fp["hi"] >= fp["lo"]      # zm = max(zf, 0): this subtree may be all sea


@pytest.mark.parametrize("d,unit_depth", [(7, 5), (8, 6)])
def test_subtree_sweep_is_invariant_under_the_partition(plb, ctx, d, unit_depth):
    """the union over the ranks of a 2- / 4-rank partition equals the 1-rank sweep: same tile count, same
    fingerprint of the per-tile statistics (sum, min, max, xor of the raw bits)"""
    import subtree_sweep as ss
    n1, fp1, plan = ss.run_sweep(plb, ctx, d, unit_depth=unit_depth)
    for world in (2, 4):
        parts = [ss.run_sweep(plb, ctx, d, rank=r, world=world, unit_depth=unit_depth) for r in range(world)]
        total = sum(p[0] for p in parts) - (world - 1) * plan.replicated_tiles()
        assert total == n1 == plan.total_tiles()
        assert np.bitwise_xor.reduce([p[1]["xor"] for p in parts]) == fp1["xor"]
        assert min(p[1]["lo"] for p in parts) == fp1["lo"] and max(p[1]["hi"] for p in parts) == fp1["hi"]
        assert abs(sum(p[1]["sum"] for p in parts) - fp1["sum"]) <= 1e-9 * abs(fp1["sum"])
    # a different unit size changes the batch shapes, not the tiles
    n2, fp2, _ = ss.run_sweep(plb, ctx, d, unit_depth=unit_depth - 2)
    assert n2 == n1 and fp2["xor"] == fp1["xor"] and fp2["lo"] == fp1["lo"] and fp2["hi"] == fp1["hi"]


def test_subtree_sweep_with_residual_params_d4(plb, ctx, oracle):
    """config 4 repeated with config-3 parameters: int16 residual tiles (a small synthetic set, reused
    by tile index) added at every swept level, flip, NEAREST storage; host-built requests"""
    import sweep
    d, W = 4, 101
    plan = sweep.SubtreeSweep(14 - d, 300, 700, d, unit_depth=4)
    amp = [0] * 11 + [5, 2.5, 1, 0.5, 0.25]
    rng = np.random.default_rng(9)
    rtiles = [rs.fractal_tile(rng, 197, 3.0) for _ in range(8)]
    rpool = ctx.pool(plb.POOL_RESID_I16, 197, len(rtiles))
    for s, t in enumerate(rtiles):
        rpool.upload(s, t)
    resid_slot = lambda level, tx, ty: (level * 7 + (tx // 2) * 3 + (ty // 2)) % len(rtiles)
    elev = ctx.pool(plb.POOL_ELEV, W, plan.capacity)
    norm = ctx.pool(plb.POOL_NORM2, W - 4, plan.capacity)
    ctx.noise_init(W)
    sc = plb.sweep_scene(noise_amp=amp, face=0, root_quad_size=100000.0, sphere=0, flip=1, elev_filter=plb.FILTER_NEAREST,
                         want_stats=0)
    sc.elev.resid_scale = 0.25
    batches = list(plan.prologue()) + [b for u in plan.units() for b in plan.unit_batches(u)]
    for level, m0, n, s0, p0, pm0 in batches:
        e, q = plb.make_requests_range(sc, level, m0, n, s0, p0, pm0)
        if level >= 10:                                       # the swept levels carry residuals
            for i in range(n):
                tx, ty = plb.morton_decode(m0 + i)
                e["resid_slot"][i] = resid_slot(level, tx, ty)
                e["rx"][i], e["ry"][i] = (tx % 2) * 96, (ty % 2) * 96
        ctx.elevation_batch(sc.elev, elev, e, resid=rpool)
        ctx.normal_batch(sc.norm, norm, elev, q)
    # the oracle takes the residual tile width (197: mod 2 windows) from an opened residual file
    geometry, _ = rs.container(min_level=0, max_level=0, tile_size=192)
    scene = oracle.make_scene(W=W, rootQuadSize=100000.0, face=0, flip=1, noiseAmp=amp, sphere=0, elev_filter=0,
                              resid=oracle.Resid(geometry))
    noise = oracle.dem_noise(W)

    def resid_of(level, tx, ty):
        if level < 10:
            return None
        return rtiles[resid_slot(level, tx, ty)].astype(np.float32) * np.float32(0.25)

    slot0, nleaf = plan.leaf_region()
    first = plan.root_morton << (2 * d)
    for m in (first, first + 37, first + nleaf - 1):
        tx, ty = plb.morton_decode(m)
        e, nrm = oracle_chain(oracle, scene, noise, 14, tx, ty, resid_of)
        s = slot0 + (m - first)
        assert np.array_equal(elev.download(s), e), (tx, ty)
        assert np.array_equal(norm.download(s), nrm), (tx, ty)


@pytest.mark.parametrize("sphere,arith", [(1, 0), (0, 0), (1, 1)])
def test_pair_batch_ids_equals_host_built_requests(plb, ctx, sphere, arith):
    """pl_pair_batch_ids (32-byte tile identities, uniforms expanded on the device) against pl_pair_batch on the
    host-built requests of the same tiles: elevation and normal tiles bit-identical; with residual tiles on the last
    level (window origin derived from tx % 2, ty % 2), distinct elevation / normal slots and a shuffled tile order"""
    amp = [-3250, -1590, -1125, -795, 561, 397] if sphere else [-140, -100, 15, 8, 5, 2.5]
    kw = dict(noise_amp=amp, face=4 if sphere else 0, root_quad_size=12720000.0 if sphere else 100000.0, sphere=sphere)
    sc = plb.sweep_scene(want_stats=1, arith=arith, **kw)
    sc.elev.resid_scale = 0.5
    L = 4
    off = [sum(4 ** k for k in range(l)) for l in range(L + 2)]
    total = off[L + 1]
    rng = np.random.default_rng(5 + sphere)
    rtiles = [rs.fractal_tile(rng, 197, 30.0) for _ in range(4)]
    rpool = ctx.pool(plb.POOL_RESID_I16, 197, len(rtiles))
    for s_, t in enumerate(rtiles):
        rpool.upload(s_, t)
    ctx.noise_init(101)
    pools = []
    for use_ids in (False, True):
        elev = ctx.pool(plb.POOL_ELEV, 101, total)
        norm = ctx.pool(plb.POOL_NORM2, 97, total + 3)
        for l in range(L + 1):
            n = 4 ** l
            perm = rng.permutation(n) if l == L else np.arange(n)
            e, q = plb.make_requests_range(sc, l, 0, n, off[l], off[l - 1] if l else 0, 0)
            ids = plb.make_tile_ids_range(l, 0, n, off[l], off[l - 1] if l else 0, 0)
            q["out_slot"] += 3                                  # the normal pool hands out other slots than the elevation pool
            ids["norm_slot"] += 3
            if l == L:
                rslot = (e["tx"] // 2 + 3 * (e["ty"] // 2)) % len(rtiles)
                e["resid_slot"], e["rx"], e["ry"] = rslot, (e["tx"] % 2) * 96, (e["ty"] % 2) * 96
                ids["resid_slot"] = rslot
            e, q, ids = e[perm], q[perm], ids[perm]
            if use_ids:
                ctx.pair_batch_ids(sc, elev, norm, ids, resid=rpool)
            else:
                ctx.pair_batch(sc.elev, sc.norm, elev, norm, e, q, resid=rpool)
        ctx.sync()
        pools.append((elev, norm))
    for s_ in list(range(0, total, 7)) + [total - 1]:
        assert np.array_equal(pools[0][0].download(s_), pools[1][0].download(s_)), s_
        assert np.array_equal(pools[0][1].download(s_ + 3), pools[1][1].download(s_ + 3)), s_
    st0 = ctx.elev_stats_range(pools[0][0], 0, total)
    st1 = ctx.elev_stats_range(pools[1][0], 0, total)
    assert np.array_equal(st0, st1)
    with pytest.raises(plb.PlError):
        bad = plb.make_tile_ids_range(2, 0, 4, 0, 0, 0)
        bad["tx"][1] = 4                                        # outside level 2
        ctx.pair_batch_ids(sc, pools[0][0], pools[0][1], bad)
    with pytest.raises(plb.PlError):
        bad = plb.make_tile_ids_range(2, 0, 4, 5, 5, 0)        # a tile that is its own parent
        ctx.pair_batch_ids(sc, pools[0][0], pools[0][1], bad)


@pytest.mark.parametrize("d,unit_depth,arith", [(6, 7, 0), (6, 7, 1), (8, 6, 0)])
def test_produce_levels_equals_the_level_by_level_sweep(plb, ctx, d, unit_depth, arith):
    """pl_produce_levels (all levels of a chain / unit in ONE launch, parent -> child dependency resolved inside the
    kernel through ready flags) against the same sweep as one pl_produce_range per level: every leaf statistic and sampled
    elevation + normal tiles bit-identical, over several repetitions (a missed dependency would read a half-written
    parent)"""
    import subtree_sweep as ss
    import sweep
    plan = sweep.SubtreeSweep(14 - d, (1 << (14 - d)) // 3, (1 << (14 - d)) // 5, d, unit_depth)
    first = plan.root_morton << (2 * d)
    want = [first, first + 4 ** d // 3, first + 4 ** d - 1]
    kw = dict(scene_kw=dict(arith=arith), unit_depth=unit_depth)
    keep0 = {"want": want}
    n0, fp0, _ = ss.run_sweep(plb, ctx, d, keep=keep0, **kw)
    for rep in range(4):
        keep1 = {"want": want}
        launches = ctx.launches
        n1, fp1, _ = ss.run_sweep(plb, ctx, d, keep=keep1, levels=True, **kw)
        assert n1 == n0 and fp1 == fp0, rep
        for m in want:
            assert np.array_equal(keep0[m][0], keep1[m][0]) and np.array_equal(keep0[m][1], keep1[m][1]), (rep, m)
        if plan.k == 0:
            assert ctx.launches - launches == 2      # one request generation + one pair launch for all 15 levels
    with pytest.raises(plb.PlError):                 # the contract: the parents of a range are the tiles of the range before
        sc = plb.sweep_scene(noise_amp=FRACTAL, face=0, root_quad_size=100000.0, sphere=0)
        e, nrm = ctx.pool(plb.POOL_ELEV, 101, 32), ctx.pool(plb.POOL_NORM2, 97, 32)
        ctx.noise_init(101)
        ctx.produce_levels(sc, e, nrm, [(0, 0, 1, 0, 0, 0), (2, 0, 16, 1, 0, 0)])


def test_elevation_seams_at_scale(plb, ctx):
    """size-independent property at a size the oracle does not reach: the 4 096 tiles of level 6 of config 1's
    terrain (pl_produce_range, fused kernel) -- every pair of neighbouring tiles holds the same zf and zm on the 5
    texel columns / rows they share (the 2-texel border convention that makes tiles self-contained,
    src/terrain/doc/overview.txt:81-88,137-142).  The RG8 normals of the vertices both tiles own agree to within one
    unorm8 step, and almost always exactly: a tile evaluates positions relative to its own origin in fp32, so the
    last bit of a component can fall on the other side of a rounding boundary (the oracle shows the same: 1 byte
    of 93 120 at level 4)."""
    amp = [-140, -100, -15, -8, 5, 2.5, 1.5, 1, 0.5, 0.25, 0.1, 0.05]
    L = 6
    off = [(4 ** l - 1) // 3 for l in range(L + 2)]
    sc = plb.sweep_scene(noise_amp=amp, face=0, root_quad_size=100000.0, sphere=0, want_stats=1)
    elev = ctx.pool(plb.POOL_ELEV, 101, off[L + 1])
    norm = ctx.pool(plb.POOL_NORM2, 97, off[L + 1])
    ctx.noise_init(101)
    for l in range(L + 1):
        ctx.produce_range(sc, elev, norm, l, 0, 4 ** l, off[l], off[l - 1] if l else 0, 0)
    ctx.sync()
    side = 1 << L
    row = None
    seams = 0
    ndiff = [0, 0]      # normal bytes compared / differing

    def normals_close(u, v):
        d = np.abs(u.astype(np.int16) - v.astype(np.int16))
        ndiff[0] += d.size
        ndiff[1] += int((d != 0).sum())
        return d.max() <= 1

    for ty in range(side):
        cur = [(elev.download(off[L] + plb.morton_encode(tx, ty)), norm.download(off[L] + plb.morton_encode(tx, ty)))
               for tx in range(side)]
        for tx in range(side - 1):
            (a, na), (b, nb) = cur[tx], cur[tx + 1]
            assert np.array_equal(a[:, 96:101, 0], b[:, 0:5, 0]) and np.array_equal(a[:, 96:101, 2], b[:, 0:5, 2]), (tx, ty)
            assert normals_close(na[:, 96], nb[:, 0]), (tx, ty)       # normal texel 96 of a tile is texel 0 of the next
            seams += 1
        if row is not None:
            for tx in range(side):
                (a, na), (b, nb) = row[tx], cur[tx]
                assert np.array_equal(a[96:101, :, 0], b[0:5, :, 0]) and np.array_equal(a[96:101, :, 2], b[0:5, :, 2]), (tx, ty)
                assert normals_close(na[96, :], nb[0, :]), (tx, ty)
                seams += 1
        row = cur
    assert seams == 2 * side * (side - 1)
    assert ndiff[1] <= 1e-3 * ndiff[0], ndiff


def test_gather_and_push_between_two_gpus():
    """tools/gather_tiles.py on 2 GPUs (skipped on a single-GPU box): the in-place NCCL all_gather of finished tiles
    and statistics, and the fused kernel's push of finished normal tiles into the peer's pool (CUDA IPC + NVLink
    stores), both bit-identical to single-GPU production"""
    import json
    import subprocess
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533",
                          os.path.join(root, "tools", "gather_tiles.py"), "--level", "5", "--reps", "2"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert out.stdout.strip().startswith("{"), "stdout must be the JSON line only"
    assert line["identical_to_single_gpu"] is True
    assert line["push_from_the_kernel"]["identical_on_every_rank"] is True
    # the multicast push (pl_pool_create_shared + pl_pool_mc_*): one store, every GPU's pool
    assert line["multicast_push_from_the_kernel"]["identical_on_every_rank"] is True


def test_shared_pool_behaves_like_a_pool_and_checks_the_multicast_order(plb, ctx, oracle):
    """a pool on the VMM allocator (pl_pool_create_shared) is an ordinary pool for production and transfers; the
    multicast entry points refuse to be called out of order (the two-GPU behaviour is in
    test_gather_and_push_between_two_gpus)"""
    norm = ctx.pool(plb.POOL_NORM2, 97, 32, shared=True)
    plain = ctx.pool(plb.POOL_NORM2, 97, 32)
    elev = ctx.pool(plb.POOL_ELEV, 101, 32)
    ctx.noise_init(101)
    sc = plb.sweep_scene(noise_amp=[-140, -100, -15], face=0, root_quad_size=100000.0, sphere=0, want_stats=1)
    for pool in (norm, plain):
        ctx.produce_range(sc, elev, pool, 0, 0, 1, 0, 0, 0)
        ctx.produce_range(sc, elev, pool, 1, 0, 4, 1, 0, 0)
        ctx.produce_range(sc, elev, pool, 2, 0, 16, 5, 1, 0)
    ctx.sync()
    for s in range(21):
        assert np.array_equal(norm.download(s), plain.download(s)), s
    with pytest.raises(plb.PlError):                      # no CUDA IPC handle for VMM memory
        norm.export()
    with pytest.raises(plb.PlError):                      # nothing bound yet
        norm.push_to_peers(2)
    with pytest.raises(plb.PlError):                      # no object yet
        norm.mc_add_device()
    with pytest.raises(plb.PlError):
        norm.mc_bind()
    with pytest.raises(plb.PlError):                      # an ordinary pool cannot join a multicast group
        plain.mc_create(2)
    with pytest.raises(plb.PlError):                      # a group of one is not a group
        norm.mc_create(1)


def test_peer_api_errors_on_one_gpu(plb, ctx):
    """pl_pool_export / attach_peers / push_to_peers argument checking (the two-GPU behaviour is in
    test_gather_and_push_between_two_gpus)"""
    norm = ctx.pool(plb.POOL_NORM2, 97, 4)
    elev = ctx.pool(plb.POOL_ELEV, 101, 4)
    h = norm.export()
    assert h.shape == (64,) and h.any()
    with pytest.raises(plb.PlError):                      # nobody attached yet
        norm.push_to_peers(True)
    norm.attach_peers(h[None, :], 0)                      # a group of one: no peers
    with pytest.raises(plb.PlError):
        norm.push_to_peers(True)
    with pytest.raises(plb.PlError):                      # self outside the group
        elev.attach_peers(h[None, :], 1)
    with pytest.raises(plb.PlError):                      # only RG8 normal pools push
        elev.push_to_peers(True)
    norm.push_to_peers(False)
    # production is unaffected
    sc = plb.sweep_scene(noise_amp=[-140, -100], face=0, root_quad_size=100000.0, sphere=0, want_stats=1)
    ctx.noise_init(101)
    ctx.produce_range(sc, elev, norm, 0, 0, 1, 0, 0, 0)
    ctx.sync()
    assert norm.download(0).any()
