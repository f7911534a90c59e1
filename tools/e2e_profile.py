"""Where the time of the e2e path goes on the GPU box: one full planet step through
bench.PlanetSweep.run_host_ids with per-launch kernel timing on, compared with the device path."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "proland-4.0_b200"))
sys.path.insert(0, ROOT)
import numpy as np
import proland_b200 as pl
import bench

os.environ["PL_E2E_PROFILE"] = "1"
with pl.Context(0) as ctx:
    sw = bench.PlanetSweep(pl, ctx, 10, want_stats=1)
    units = sw.units
    sw.run_host_ids(units[:2]); ctx.sync()
    for name, fn in (("device path", lambda: sw.run_device(units)), ("e2e path", lambda: sw.run_host_ids(units))):
        ctx.timing_collect(); ctx.timing_enable(True)
        t0 = time.perf_counter(); fn(); ctx.sync(); dt = time.perf_counter() - t0
        kt = ctx.timing_collect(); ctx.timing_enable(False)
        print("%s: wall %.3f s; kernels: %s" % (name, dt, {k: (round(v[0], 1), v[1]) for k, v in kt.items() if v[1]}))
