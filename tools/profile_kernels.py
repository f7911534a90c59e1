"""Profiling workload for ncu: levels 0..7 of one planet face as whole-level batches,
then 16384 level-8 tiles (positive noise amplitude -> the slope/curvature path, sphere
normals) as ONE batch: launch #9 of each kernel is the one to capture.

  ncu --set full --clock-control none --import-source on -k regex:elevation_kernel -s 8 -c 1 \
      -o gpurun_out/elev python tools/profile_kernels.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "proland-4.0_b200"))
import proland_b200 as pl

AMP = [-3250, -1590, -1125, -795, -561, -397, -140, -100, 15, 8, 5, 2.5, 1.5, 1, 0.5, 0.25, 0.1, 0.05]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
with pl.Context(0) as ctx:
    arith = pl.ARITH_FAST if os.environ.get("PL_ARITH", "exact") == "fast" else pl.ARITH_EXACT
    sc = pl.sweep_scene(noise_amp=AMP, face=3, root_quad_size=12720000.0, sphere=1, want_stats=1, arith=arith)
    off = [sum(4 ** k for k in range(l)) for l in range(10)]
    elev = ctx.pool(pl.POOL_ELEV, 101, off[8] + 16384)
    norm = ctx.pool(pl.POOL_NORM2, 97, off[8] + 16384)
    ctx.noise_init(101)
    if os.environ.get("PL_NO_FUSE"):
        ctx.no_fuse(True)      # time the elevation and the normal kernel separately
    for l in range(8):
        ctx.produce_range(sc, elev, norm, l, 0, 4 ** l, off[l], off[l - 1] if l else 0, 0)
    ctx.timing_enable(True)
    for _ in range(reps):
        ctx.produce_range(sc, elev, norm, 8, 0, 16384, off[8], off[7], 0)
    print(ctx.timing_collect())
