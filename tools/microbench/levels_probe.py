import sys, os
sys.path.insert(0, "/root/repo/proland-4.0_b200")
import proland_b200 as pl
PLANET = [-3250, -1590, -1125, -795, -561, -397, -140, -100, 15, 8, 5, 2.5]
with pl.Context(0) as ctx:
    n = 16384                       # level 7
    o4, o5, o6, o7 = 0, 256, 256 + 1024, 256 + 1024 + 4096
    elev = ctx.pool(pl.POOL_ELEV, 101, o7 + n)
    norm = ctx.pool(pl.POOL_NORM2, 97, o7 + n)
    ctx.noise_init(101)
    sc = pl.sweep_scene(noise_amp=PLANET, face=3, root_quad_size=12720000.0, sphere=1, want_stats=1)
    ctx.produce_range(sc, elev, norm, 5, 0, 1024, o5, o4, 0)
    ctx.produce_range(sc, elev, norm, 6, 0, 4096, o6, o5, 0)
    for name, fn in (("range", lambda: ctx.produce_range(sc, elev, norm, 7, 0, n, o7, o6, 0)),
                     ("levels-1", lambda: ctx.produce_levels(sc, elev, norm, [(7, 0, n, o7, o6, 0)])),
                     ("levels-2", lambda: ctx.produce_levels(sc, elev, norm, [(6, 0, n // 4, o6, o5, 0), (7, 0, n, o7, o6, 0)])),
                     ("range-2", lambda: (ctx.produce_range(sc, elev, norm, 6, 0, n // 4, o6, o5, 0), ctx.produce_range(sc, elev, norm, 7, 0, n, o7, o6, 0)))):
        fn(); ctx.sync(); ctx.timing_collect(); ctx.timing_enable(True)
        for _ in range(5): fn()
        t = ctx.timing_collect(); ctx.timing_enable(False)
        print(name, "pair ms per call", t["pair"][0] / 5, "gen ms", t["genreq"][0] / 5)
    # a whole chain: levels 0..7 of the face (21 845 tiles) as one launch vs eight
    off = [sum(4 ** k for k in range(l)) for l in range(9)]
    elev2 = ctx.pool(pl.POOL_ELEV, 101, off[8])
    norm2 = ctx.pool(pl.POOL_NORM2, 97, off[8])
    rr = [(l, 0, 4 ** l, off[l], off[l - 1] if l else 0, 0) for l in range(8)]
    import time
    for name, fn in (("range-8", lambda: [ctx.produce_range(sc, elev2, norm2, *r) for r in rr]),
                     ("levels-8", lambda: ctx.produce_levels(sc, elev2, norm2, rr)),
                     ("levels-6+2", lambda: (ctx.produce_levels(sc, elev2, norm2, rr[:6]), ctx.produce_levels(sc, elev2, norm2, rr[6:])))):
        fn(); ctx.sync(); ctx.timing_collect(); ctx.timing_enable(True)
        t0 = time.perf_counter()
        for _ in range(5): fn()
        ctx.sync()
        wall = (time.perf_counter() - t0) / 5
        t = ctx.timing_collect(); ctx.timing_enable(False)
        print(name, "pair ms per call", t["pair"][0] / 5, "gen ms", t["genreq"][0] / 5, "wall ms", wall * 1e3)
