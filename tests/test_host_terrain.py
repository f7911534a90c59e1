"""CPU tests of the caller side of the path: TerrainNode/TerrainQuad subdivision (reference:
core/sources/proland/terrain/TerrainQuad.cpp:81-173, TerrainNode.cpp:92-145) and TileSampler's
per-frame put / get / prefetch of tiles (TileSampler.cpp:304-496), with the recording producer."""
import math

import pytest


@pytest.fixture(scope="module")
def ph():
    import proland_host
    proland_host.build()
    return proland_host


def quadtree_restatement(size, cam, split_dist, max_level, ground=0.0):
    """the split rule, restated: split iff dist < l * splitDist and level < maxLevel;
    dist = max(|cz - zmax|, max(min |cx - xmin|,|cx - xmax|, min |cy - ymin|,|cy - ymax|)) in float"""
    import numpy as np
    out = []

    def visit(level, tx, ty, ox, oy, l):
        zmax = max(0.0, ground)
        d = max(abs(cam[2] - zmax), max(min(abs(cam[0] - ox), abs(cam[0] - (ox + l))), min(abs(cam[1] - oy), abs(cam[1] - (oy + l)))))
        split = np.float32(d) < np.float32(np.float32(l) * np.float32(split_dist)) and level < max_level
        out.append((level, tx, ty, 0 if split else 1))
        if split:
            hl = float(np.float32(l) / np.float32(2.0))
            visit(level + 1, 2 * tx, 2 * ty, ox, oy, hl)
            visit(level + 1, 2 * tx + 1, 2 * ty, ox + hl, oy, hl)
            visit(level + 1, 2 * tx, 2 * ty + 1, ox, oy + hl, hl)
            visit(level + 1, 2 * tx + 1, 2 * ty + 1, ox + hl, oy + hl, hl)

    visit(0, 0, 0, -size, -size, 2.0 * size)
    return out


def test_split_distance(ph):
    f = ph.lib().plh_split_distance
    want = 2.0 * 1024.0 / 1024.0 * math.tan(math.radians(40.0)) / math.tan(math.radians(80.0) / 2.0)
    assert abs(f(2.0, 1024.0, math.radians(80.0)) - want) < 1e-5
    assert abs(f(2.0, 1920.0, math.radians(60.0)) - 2.0 * 1920 / 1024 * math.tan(math.radians(40)) / math.tan(math.radians(30))) < 1e-4
    assert f(0.1, 1024.0, math.radians(80.0)) == pytest.approx(1.1)      # never below 1.1 (TerrainNode.cpp:130-132)


@pytest.mark.parametrize("cam", [(0.0, 0.0, 100000.0), (1234.5, -20000.25, 500.0), (-49999.0, 49999.0, 10.0),
                                 (70000.0, 3.0, 2000.0), (12.0, 7.0, 1.0)])
def test_quadtree_matches_the_split_rule(ph, cam):
    t = ph.Terrain(50000.0, split_factor=2.0, max_level=12)
    sd = ph.lib().plh_split_distance(2.0, 1024.0, math.radians(80.0))
    n = t.update(*cam, split_dist=sd)
    quads = t.quads()
    assert n == len(quads)
    assert quads == quadtree_restatement(50000.0, cam, sd, 12)
    # moving away merges quads again; coming back restores the same tree
    t.update(cam[0], cam[1], 1e7, split_dist=sd)
    assert t.quads() == [(0, 0, 0, 1)]
    t.update(*cam, split_dist=sd)
    assert t.quads() == quads
    t.close()


def test_sampler_holds_exactly_the_tiles_of_the_quads(ph):
    s = ph.TestScene(capacity=512, max_level=30)
    t = ph.Terrain(50000.0, max_level=8)
    sm = ph.Sampler("elevationSampler", s.producer)
    sd = 1.4
    path = [(-30000.0 + 1500.0 * k, 10000.0 * math.sin(k / 5.0), 800.0) for k in range(40)]
    made = 0
    for cam in path:
        t.update(*cam, split_dist=sd)
        quads = t.quads()
        ph.frame_update(s.scheduler, t, [sm])
        # storeLeaf + storeParent: one tile per quad, all produced, nothing else in use
        assert sm.tile_count == len(quads)
        assert s.cache.stats()["used"] == len(quads)
        for (level, tx, ty, leaf) in quads:
            tile = s.producer.find_tile(level, tx, ty)
            assert tile is not None and tile.done
        # the root quad size reached the producer (TileSampler.cpp:419-421)
        assert s.producer.info()["id"] == 0
        made = len(s.calls())
    assert made > 100
    # every tile was produced after its parent
    seen = set()
    for (level, tx, ty, _) in s.calls():
        assert level == 0 or (level - 1, tx // 2, ty // 2) in seen
        seen.add((level, tx, ty))
    # every production is a cache miss (tiles still cached are never made again; evicted ones are)
    assert len(s.calls()) == s.cache.stats()["misses"] and len(seen) <= len(s.calls())
    sm.close()
    assert s.cache.stats()["used"] == 0
    t.close()
    s.close()


def test_sampler_store_parent_false_keeps_only_leaves(ph):
    s = ph.TestScene(capacity=2048, max_level=30)
    t = ph.Terrain(50000.0, max_level=6)
    sm = ph.Sampler("leafSampler", s.producer, store_parent=False)
    t.update(100.0, 200.0, 50.0, split_dist=1.5)
    ph.frame_update(s.scheduler, t, [sm])
    leaves = [q for q in t.quads() if q[3]]
    assert sm.tile_count == len(leaves)
    for (level, tx, ty, _) in leaves:
        assert s.producer.find_tile(level, tx, ty).done
    sm.close()
    t.close()
    s.close()


def test_sync_sampler_prefetches_children_of_new_leaves(ph):
    s = ph.TestScene(capacity=256, max_level=30, prefetch_rate=8, prefetch_queue=64)
    t = ph.Terrain(50000.0, max_level=3)
    sm = ph.Sampler("s", s.producer)
    t.update(0.0, 0.0, 100.0, split_dist=1.2)
    ph.frame_update(s.scheduler, t, [sm])
    n_quads = len(t.quads())
    assert s.cache.stats()["used"] == n_quads and s.scheduler.stats()["queued"] == 0
    # second frame: the leaves are new trees -> their four children are prefetched (TileSampler.cpp:463-496)
    ph.frame_update(s.scheduler, t, [sm])
    st = s.cache.stats()
    assert st["used"] == n_quads and st["unused"] > 0
    leaves = [q for q in t.quads() if q[3]]
    # ... as far as spare capacity goes: prefetchCount = unused + free slots (TileSampler.cpp:313)
    assert st["unused"] == min(4 * len(leaves), 256 - n_quads)
    # ... and produced at prefetchRate tasks per frame
    done_before = len(s.calls())
    ph.frame_update(s.scheduler, t, [sm])
    assert 0 < len(s.calls()) - done_before <= 8 * 4          # 8 task graphs, each may pull in ancestors
    sm.close()
    t.close()
    s.close()


def test_async_sampler_takes_only_cached_tiles_and_prefetches_the_rest(ph):
    """earth-srtm-async (TileSampler.cpp:430-441): below the root a tile is taken only if the cache
    already knows it (findTile with includeCache), a missing LEAF is prefetched instead; a frame later
    the prefetched tiles are in the cache and are taken -- with their tasks, if those have not run yet"""
    s = ph.TestScene(capacity=256, max_level=30, prefetch_rate=2, prefetch_queue=64)
    t = ph.Terrain(50000.0, max_level=4)
    sm = ph.Sampler("s", s.producer, asynchronous=True)
    t.update(20000.0, -20000.0, 100.0, split_dist=1.3)
    quads = t.quads()
    leaves = [q for q in quads if q[3]]
    ph.frame_update(s.scheduler, t, [sm])
    # only the root is synchronous; the scheduler also started prefetchRate = 2 of the queued leaf tasks
    made0 = [c[:3] for c in s.calls()]
    assert sm.tile_count == 1 and made0[0] == (0, 0, 0) and len(made0) <= 1 + 2 * 4
    assert sum(1 for c in made0 if (c[0], c[1], c[2], 1) in leaves) == 2
    st = s.cache.stats()
    # every leaf was prefetched: it sits in the cache unused, its ancestors pinned by its task
    for (level, tx, ty, _) in leaves:
        assert s.producer.find_tile(level, tx, ty, include_cache=True) is not None
    assert st["used"] + st["unused"] == len(quads)
    assert s.scheduler.stats()["queued"] == min(len(leaves), 64) - 2
    ph.frame_update(s.scheduler, t, [sm])
    assert sm.tile_count == len(quads) and len(s.calls()) == len(quads)
    for (level, tx, ty, _) in quads:
        assert s.producer.find_tile(level, tx, ty).done
    sm.close()
    t.close()
    s.close()
