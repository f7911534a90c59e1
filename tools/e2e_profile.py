"""Where the host time of the e2e path goes: cumulative wall time of request building (per thread
count), pl_pair_batch (validation + staging copy + launch) and the statistics read-back, for one
level-2 subtree of the planet sweep (levels 3..10, 87 380 pairs)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "proland-4.0_b200"))
sys.path.insert(0, ROOT)
import numpy as np
import proland_b200 as pl
import bench

with pl.Context(0) as ctx:
    sw = bench.PlanetSweep(pl, ctx, 10, want_stats=1)
    units = sw.units[:2]
    todo = list(sw.plan.batches(units, 10))
    bufs = (np.zeros(65536, pl.ELEV_REQ_DTYPE), np.zeros(65536, pl.NORM_REQ_DTYPE))
    for nt in (1, 4, 16, 0):
        t_gen = t_sub = t_stat = 0.0
        n_tot = 0
        for f, level, m0, n, s0, p0, pm0 in todo:
            sc = sw.scenes[f]
            t0 = time.perf_counter()
            e, q = pl.make_requests_range(sc, level, m0, n, s0, p0, pm0, nthreads=nt, out=bufs)
            t1 = time.perf_counter()
            ctx.pair_batch(sc.elev, sc.norm, sw.elev, sw.norm, e, q)
            t2 = time.perf_counter()
            if n >= 4096:
                ctx.elev_stats_range(sw.elev, s0, n)
            t3 = time.perf_counter()
            t_gen += t1 - t0; t_sub += t2 - t1; t_stat += t3 - t2; n_tot += n
        print("threads=%2d  %d pairs: build %.1f ms (%.0f ns/tile), pair_batch %.1f ms (%.0f ns/tile), stats+wait %.1f ms"
              % (nt, n_tot, t_gen * 1e3, t_gen / n_tot * 1e9, t_sub * 1e3, t_sub / n_tot * 1e9, t_stat * 1e3))
    print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
