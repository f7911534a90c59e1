"""CPU tests of the C++ host layer (proland-4.0_b200/host): TileStorage / TileCache / TileProducer /
task graphs / BatchScheduler semantics against the reference's contract
(producer/TileCache.cpp:150-336, TileProducer.cpp:44-353,709-791), with a recording producer
instead of device work, plus the XML resource surface.  No GPU needed."""
import os
import re

import pytest


@pytest.fixture(scope="module")
def ph():
    import proland_host
    proland_host.build()
    return proland_host


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_library_exports_every_declared_symbol(ph):
    hdr = open(os.path.join(ROOT, "proland-4.0_b200", "host", "proland_host.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(plh_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(ph.EXPORTS), declared ^ set(ph.EXPORTS)
    for name in declared:
        assert hasattr(ph.lib(), name), name


def test_get_tile_builds_the_parent_chain_and_runs_it_in_waves(ph):
    s = ph.TestScene(capacity=8)
    p = s.producer
    assert p.task_type == "CreateRecordedTile" and p.info()["border"] == 0 and not p.info()["gpu"]
    t = p.get_tile(3, 5, 2)
    # startCreateTile acquired the whole ancestor chain: 4 used tiles, nothing produced yet
    assert s.cache.stats() == dict(used=4, unused=0, capacity=8, free=4, queries=4, misses=4)
    assert not t.done and s.calls() == []
    s.scheduler.run([t])
    # parents first, one wave per level (each tile depends on its parent)
    assert [c[:3] for c in s.calls()] == [(0, 0, 0), (1, 1, 0), (2, 2, 1), (3, 5, 2)]
    assert s.scheduler.stats()["waves"] == 4 and s.begin_end() == (4, 4)
    # stopCreateTile released the inputs: ancestors are unused (cached), the requested tile stays used
    assert t.done and s.cache.stats()["used"] == 1 and s.cache.stats()["unused"] == 3
    # siblings share the chain: only the new leaf is produced, in one wave
    t2 = p.get_tile(3, 4, 2)
    s.scheduler.run([t2])
    assert [c[:3] for c in s.calls()[4:]] == [(3, 4, 2)]
    assert s.scheduler.stats()["waves"] == 5
    p.put_tile(t)
    p.put_tile(t2)
    s.close()


def test_independent_tiles_of_one_level_share_a_wave(ph):
    s = ph.TestScene(capacity=64)
    tiles = [s.producer.get_tile(2, tx, ty) for ty in range(4) for tx in range(4)]
    s.scheduler.run(tiles)
    st = s.scheduler.stats()
    assert st["waves"] == 3 and st["tasks"] == 1 + 4 + 16
    levels = [c[0] for c in s.calls()]
    assert levels == sorted(levels)
    for t in tiles:
        s.producer.put_tile(t)
    s.close()


def test_users_count_and_find_tile(ph):
    s = ph.TestScene(capacity=4)
    p = s.producer
    a = p.get_tile(0, 0, 0)
    b = p.get_tile(0, 0, 0)                       # second user of the same tile: no new query
    assert a.h == b.h and s.cache.stats()["queries"] == 1
    s.scheduler.run([a])
    p.put_tile(a)
    assert s.cache.stats()["used"] == 1           # still one user
    assert p.find_tile(0, 0, 0) is not None
    p.put_tile(b)
    assert s.cache.stats()["used"] == 0 and s.cache.stats()["unused"] == 1
    # unused tiles are found only with includeCache (TileCache.cpp:166-172)
    assert p.find_tile(0, 0, 0) is None
    assert p.find_tile(0, 0, 0, include_cache=True) is not None
    # ... and come back without being produced again
    c = p.get_tile(0, 0, 0)
    assert c.done and len(s.calls()) == 1 and s.cache.stats()["misses"] == 1
    p.put_tile(c)
    s.close()


def test_lru_eviction_order_and_cache_full(ph):
    s = ph.TestScene(capacity=3, max_level=0)
    p = s.producer
    # three independent root tiles of three "faces" (tx distinguishes them; level 0 has no parent)
    tiles = [p.get_tile(0, k, 0) for k in range(3)]
    s.scheduler.run(tiles)
    with pytest.raises(ph.HostError, match="Insufficient tile cache size"):
        p.get_tile(0, 3, 0)                        # all three slots are in use -> NULL in the reference
    for k in (1, 0, 2):                            # released in this order: 1 is the least recently used
        p.put_tile(tiles[k])
    d = p.get_tile(0, 3, 0)                        # evicts tile 1, reuses its slot
    s.scheduler.run([d])
    slot_of_1 = [c[3] for c in s.calls() if c[:3] == (0, 1, 0)][0]      # (order inside a wave is free)
    assert s.calls()[-1][:3] == (0, 3, 0) and s.calls()[-1][3] == slot_of_1
    assert p.find_tile(0, 1, 0, include_cache=True) is None
    assert p.find_tile(0, 0, 0, include_cache=True) is not None
    # touching tile 0 makes it the most recently used: the next eviction takes tile 2
    t0 = p.get_tile(0, 0, 0)
    p.put_tile(t0)
    e = p.get_tile(0, 4, 0)
    assert p.find_tile(0, 2, 0, include_cache=True) is None
    assert p.find_tile(0, 0, 0, include_cache=True) is not None
    s.scheduler.run([e])
    p.put_tile(d)
    p.put_tile(e)
    s.close()


def test_evicted_tile_is_produced_again_on_request(ph):
    s = ph.TestScene(capacity=2, max_level=0)
    p = s.producer
    a = p.get_tile(0, 0, 0)
    s.scheduler.run([a])
    p.put_tile(a)
    b, c = p.get_tile(0, 1, 0), p.get_tile(0, 2, 0)     # c evicts a
    s.scheduler.run([b, c])
    p.put_tile(b)
    p.put_tile(c)
    a2 = p.get_tile(0, 0, 0)                             # data gone: must run again
    assert not a2.done
    s.scheduler.run([a2])
    got = [x[:3] for x in s.calls()]
    assert got[0] == got[3] == (0, 0, 0) and sorted(got[1:3]) == [(0, 1, 0), (0, 2, 0)]   # order inside a wave is free
    p.put_tile(a2)
    s.close()


def test_invalidate_reruns_the_tile_and_what_depends_on_it(ph):
    s = ph.TestScene(capacity=8)
    p = s.producer
    t = p.get_tile(2, 3, 3)
    s.scheduler.run([t])
    n0 = len(s.calls())
    s.scheduler.run([t])                                  # nothing to do
    assert len(s.calls()) == n0
    p.invalidate_tile(1, 1, 1)                            # the parent's data changed
    assert not t.done                                     # its graph holds the parent's task: not done any more
    s.scheduler.run([t])                                  # the parent runs again, which makes the child stale
    assert [c[:3] for c in s.calls()[n0:]] == [(1, 1, 1), (2, 3, 3)]
    p.invalidate_tiles()
    s.scheduler.run([t])
    assert [c[:3] for c in s.calls()[n0 + 2:]] == [(0, 0, 0), (1, 1, 1), (2, 3, 3)]
    p.put_tile(t)
    s.close()


def test_prefetch_queue_and_rate(ph):
    s = ph.TestScene(capacity=16, max_level=0, prefetch_rate=2, prefetch_queue=3)
    p = s.producer
    for k in range(4):
        assert p.prefetch_tile(0, k, 0)
    assert not p.prefetch_tile(0, 0, 0)                   # already in the cache: nothing to do
    assert s.scheduler.stats()["queued"] == 3             # queue of 3: the oldest request was dropped
    assert s.cache.stats()["unused"] == 4                 # prefetched tiles start unused (TileCache.cpp:283-285)
    s.scheduler.run([])
    assert len(s.calls()) == 2                            # prefetchRate tasks per run
    s.scheduler.run([])
    assert len(s.calls()) == 3 and s.scheduler.stats()["queued"] == 0
    # without prefetch support prefetchTile declines (TileProducer.cpp:521-532)
    s2 = ph.TestScene(capacity=4)
    assert not s2.producer.prefetch_tile(0, 0, 0)
    s.close()
    s2.close()


def test_has_children_and_upsample_variants(ph):
    s = ph.TestScene(capacity=2, max_level=3)
    assert s.producer.has_tile(3, 0, 0) and not s.producer.has_tile(4, 0, 0)
    assert s.producer.has_children(2, 1, 1) and not s.producer.has_children(3, 1, 1)
    s.close()
    assert ph.upsample_variant("upsampleShader;") == (True, False)
    assert ph.upsample_variant("upsampleShader-noClamp;") == (True, True)
    assert ph.upsample_variant("upsampleShader-plain;") == (False, False)
    assert ph.upsample_variant("blendShader;") is None


ARCHIVE = """<!-- a licence comment before the declaration, like the reference's archives -->
<?xml version="1.0" ?>
<archive>
    <multithreadScheduler name="defaultScheduler" nthreads="3" fps="0"/>
    <tileCache name="groundElevations" scheduler="defaultScheduler">
        <gpuTileStorage tileSize="101" nTiles="512"
            internalformat="RGB32F" format="RGB" type="FLOAT" min="LINEAR" mag="LINEAR"/>
    </tileCache>
    <elevationProducer name="groundElevations1" cache="groundElevations"
        noise="-140,-100,-15,-8,5,2.5,1.5,1,0.5,0.25,0.1,0.05"/>
    <tileCache name="groundNormals" scheduler="defaultScheduler">
        <gpuTileStorage tileSize="97" nTiles="512"
            internalformat="RG8" format="RG" type="FLOAT" min="LINEAR" mag="LINEAR"/>
    </tileCache>
    <normalProducer name="groundNormals1" cache="groundNormals" elevations="groundElevations1"/>
    <terrainNode name="terrain" size="50000" zmin="0" zmax="5000" splitFactor="2" maxLevel="16"/>
</archive>
"""


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="CPU-only behaviour")
def test_gpu_resources_fail_loudly_without_a_device(ph):
    """no CPU fallback: creating a gpuTileStorage without a CUDA device is an error"""
    ph.lib().plh_quiet_errors(1)
    with ph.Scene(ARCHIVE) as scene:
        with pytest.raises(ph.HostError):
            scene.producer("groundNormals1")
        # the scheduler is plain host code
        assert scene.scheduler("defaultScheduler").stats()["frame"] == 0
    ph.lib().plh_quiet_errors(0)


def test_archive_errors(ph):
    ph.lib().plh_quiet_errors(1)
    with pytest.raises(ph.HostError, match="closes"):
        ph.Scene("<archive><tileCache name='a'></archive>")
    with ph.Scene(ARCHIVE) as scene:
        with pytest.raises(ph.HostError, match="Missing or invalid resource"):
            scene.producer("nope")
        with pytest.raises(ph.HostError, match="not a TileProducer"):
            scene.producer("defaultScheduler")
    # unknown attributes are rejected like Resource::checkParameters does
    bad = ARCHIVE.replace('nthreads="3"', 'nthreads="3" colour="red"')
    with ph.Scene(bad) as scene:
        with pytest.raises(ph.HostError, match="unsupported 'colour' attribute"):
            scene.scheduler("defaultScheduler")
    with pytest.raises(ph.HostError, match="duplicate resource name"):
        ph.Scene(ARCHIVE.replace('name="groundNormals1" cache', 'name="groundElevations1" cache'))
    ph.lib().plh_quiet_errors(0)


def test_cache_full_while_acquiring_inputs_leaves_the_cache_consistent(ph):
    """the reference asserts when an input tile cannot be acquired; here CacheFullError, and every slot
    and user count taken on the way is given back"""
    ph.lib().plh_quiet_errors(1)
    s = ph.TestScene(capacity=2)
    with pytest.raises(ph.HostError, match="Insufficient tile cache size"):
        s.producer.get_tile(3, 1, 1)                 # needs 4 slots
    assert s.cache.stats()["used"] == 0 and s.cache.stats()["unused"] == 0 and s.cache.stats()["free"] == 2
    t = s.producer.get_tile(1, 1, 1)                 # 2 slots: fits
    s.scheduler.run([t])
    assert [c[:3] for c in s.calls()] == [(0, 0, 0), (1, 1, 1)]
    s.producer.put_tile(t)
    s.close()
    ph.lib().plh_quiet_errors(0)
