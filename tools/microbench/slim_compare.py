import sys, os
sys.path.insert(0, "proland-4.0_b200")
import proland_b200 as pl
PLANET = [-3250, -1590, -1125, -795, -561, -397, -140, -100, 15, 8, 5, 2.5, 1.5, 1, 0.5, 0.25, 0.1, 0.05]
SRTM = [0] * 7 + [5, 2.5, 1, 0.5]
level, reps = 8, 20
n = 4 ** 7
with pl.Context(0) as ctx:
    elev = ctx.pool(pl.POOL_ELEV, 101, n + n // 4 + 1); norm = ctx.pool(pl.POOL_NORM2, 97, n + n // 4 + 1); ctx.noise_init(101)
    for name, kw in (("planet LINEAR", dict(noise_amp=PLANET, face=3, root_quad_size=12720000.0, sphere=1)),
                     ("srtm NEAREST flip", dict(noise_amp=SRTM, face=2, root_quad_size=12720000.0, sphere=1, flip=1, elev_filter=pl.FILTER_NEAREST)),
                     ("flat LINEAR", dict(noise_amp=PLANET[6:], face=0, root_quad_size=100000.0, sphere=0))):
        sc = pl.sweep_scene(want_stats=1, arith=pl.ARITH_FAST, **kw)
        out = []
        for rnd in range(2):
            for mode in (1, 0):
                ctx.no_slim(mode)
                ctx.produce_range(sc, elev, norm, level - 1, 0, n // 4, n, 0, 0)
                for _ in range(3): ctx.produce_range(sc, elev, norm, level, 0, n, 0, n, 0)
                ctx.sync(); ctx.timing_collect(); ctx.timing_enable(True)
                for _ in range(reps): ctx.produce_range(sc, elev, norm, level, 0, n, 0, n, 0)
                ms, cnt, tiles = ctx.timing_collect()["pair"]; ctx.timing_enable(False)
                out.append(("regular" if mode else "slim", round(ms / cnt, 4)))
        print(name, out)
