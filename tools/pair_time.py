"""Kernel-only time of the fused elevation + normal kernel on one full level of a planet face, both arithmetic
contracts (PL_ARITH_EXACT / PL_ARITH_FAST): CUDA events of the library (pl_timing_*), L2 flushed by size (a level-8
launch writes 2.3 GB).  python tools/pair_time.py [level] [reps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "proland-4.0_b200"))
import proland_b200 as pl  # noqa: E402

PLANET = [-3250, -1590, -1125, -795, -561, -397, -140, -100, 15, 8, 5, 2.5, 1.5, 1, 0.5, 0.25, 0.1, 0.05]
level = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
n = 4 ** min(level, 7)
peak = 6554.2
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
out = {}
with pl.Context(0) as ctx:
    elev = ctx.pool(pl.POOL_ELEV, 101, n + n // 4 + 1)
    norm = ctx.pool(pl.POOL_NORM2, 97, n + n // 4 + 1)
    ctx.noise_init(101)
    for name, kw in (("planet", dict(noise_amp=PLANET, face=3, root_quad_size=12720000.0, sphere=1)),
                     ("flat", dict(noise_amp=PLANET[6:], face=0, root_quad_size=100000.0, sphere=0))):
        for arith in (pl.ARITH_EXACT, pl.ARITH_FAST):
            sc = pl.sweep_scene(want_stats=1, arith=arith, **kw)
            # parents: one level-(level-1) range produced from garbage parents is fine for timing
            ctx.produce_range(sc, elev, norm, level - 1, 0, n // 4, n, 0, 0)
            for _ in range(3):
                ctx.produce_range(sc, elev, norm, level, 0, n, 0, n, 0)
            ctx.sync()
            ctx.timing_collect()
            ctx.timing_enable(True)
            for _ in range(reps):
                ctx.produce_range(sc, elev, norm, level, 0, n, 0, n, 0)
            ms, cnt, tiles = ctx.timing_collect()["pair"]
            ctx.timing_enable(False)
            per = ms / cnt
            gbs = 192098 * n / (per * 1e-3) / 1e9
            out["%s_%s" % (name, "fast" if arith else "exact")] = {"ms_per_launch": per, "tiles": n, "ns_per_pair": per * 1e6 / n,
                                                                  "algorithmic_gbs": gbs, "frac_of_peak": gbs / peak}
print(json.dumps(out, indent=1))
