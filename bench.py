#!/usr/bin/env python
"""bench.py -- elevation+normal tile pairs/sec on B200 (BASELINE.json's metric).

Workload (config.workload): demo-fractalplanet -- the six cube faces of the
fractal planet, every quadtree tile of levels 0..10 (8 388 606 elevation+normal
pairs, sphere-deformed RG8 normals, slope/curvature-modulated noise).  One
"step" produces the whole planet once.  The planet is cut into 96 subtrees (one
per level-2 quad of each face); rank r owns subtrees r, r+N, ... and produces
them breadth-first into a recycling device pool (no data-path collective; NCCL
only gathers per-rank counters).  Total work is fixed as N grows -> "strong".

  value : device-resident path -- pl_produce_range: the per-tile uniforms are
          generated on the GPU, nothing but (level, Morton range) crosses PCIe.
  e2e   : the per-tile plugin path -- every tile handed over by identity
          (level, tx, ty and its slots: pl_tile_id, 32 bytes) as HOST arrays
          through pl_pair_batch_ids (copied to the device inside the timed
          region, uniforms expanded there), per-tile (zmin,zmax) read back.

  python bench.py [--gpus N --steps K --warmup W] [--impl reference]
  torchrun ... bench.py --gpus N ...          (one rank per GPU)
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "proland-4.0_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

PLANET_AMP = [-3250, -1590, -1125, -795, -561, -397, -140, -100, 15, 8, 5, 2.5, 1.5, 1, 0.5, 0.25,
              0.1, 0.05]
PLANET_SIZE = 12720000.0   # 2 * R, R = 6 360 000 m (fractalplanet.xml)
ELEV_BYTES = 101 * 101 * 12 + 54 * 54 * 4        # elevation kernel: write + parent window
ELEV_BYTES_L0 = 101 * 101 * 12
NORM_BYTES = 99 * 99 * 4 + 97 * 97 * 2           # normal kernel: own zm read + RG8 write
PAIR_BYTES = ELEV_BYTES + NORM_BYTES             # 192 098
METRIC = "elevation+normal tile pairs/sec"
# dram__bytes_read.sum + dram__bytes_write.sum per tile of one `ncu --set full` capture (profiles/README.md):
# 16384 level-8 planet tiles in one launch.  pair: the kernel of the timed contract -- PL_ARITH_FAST
# profiles/pair_r2x_fast_ncu_raw.csv (233.36 MB read + 2 361.59 MB written), PL_ARITH_EXACT profiles/pair_r1w_ncu_raw.csv;
# the residual variants (config 3) and the flat scene (config 1) have no capture of their own: their records say null
TRAFFIC = {"pair": 158382, "elevation": 136520, "normal": 59565}
TRAFFIC_EXACT_PAIR = 155752
ARITH_NOTE = {"fast": "PL_ARITH_FAST: elevation tiles bit-identical to the oracle; a normal byte within ONE unorm8 step of "
                      "the canonical evaluation (tests/test_gpu_fast.py bounds how many differ by the reference's own "
                      "non-contracted reading)",
              "exact": "PL_ARITH_EXACT: elevation and normal tiles bit-identical to the oracle"}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------- the sweep

class PlanetSweep:
    """Breadth-first production of level-2 subtrees into a recycling pool (plan: sweep.py)."""

    def __init__(self, pl, ctx, max_level, want_stats=1, arith=0):
        import sweep as plan
        self.plan, self.pl, self.ctx, self.max_level = plan, pl, ctx, max_level
        _, self.capacity = plan.region_offsets(max_level)
        self.elev = ctx.pool(pl.POOL_ELEV, 101, self.capacity)
        self.norm = ctx.pool(pl.POOL_NORM2, 97, self.capacity)
        ctx.noise_init(101)
        self.want_stats = want_stats
        self.units = plan.planet_units()
        self.set_arith(arith)

    def set_arith(self, arith):
        """the normal pass's arithmetic contract (pl_norm_scene.arith): elevations are bit-exact under both"""
        pl = self.pl
        self.arith = arith
        self.scenes = {f: pl.sweep_scene(noise_amp=PLANET_AMP, face=f, root_quad_size=PLANET_SIZE,
                                         sphere=1, elev_filter=pl.FILTER_LINEAR,
                                         want_stats=self.want_stats, arith=arith) for f in range(1, 7)}

    def pairs_of(self, units, count_roots):
        return self.plan.pairs_in_units(units, self.max_level, count_roots)

    def run_device(self, units):
        pr = self.ctx.produce_range
        for f, level, m0, n, s0, p0, pm0 in self.plan.batches(units, self.max_level):
            pr(self.scenes[f], self.elev, self.norm, level, m0, n, s0, p0, pm0)

    def run_host_ids(self, units):
        """e2e: what a caller of the plugin path holds per tile -- its quadtree coordinates and the slots its caches
        handed out, 32 bytes (pl_tile_id) -- as HOST arrays through pl_pair_batch_ids: validated, copied into the
        pinned staging ring, uploaded on the copy stream beside the kernels of earlier batches; the uniforms of
        doCreateTile (noise layer / rotation, windows, fp64 patch geometry) are expanded on the device.  Per-tile
        (zmin, zmax) come back through the asynchronous read-back and are collected three read-backs late -- the
        reference's TileSamplerZ collects its read-backs a few frames late in the same way (ReadbackManager).
        Nothing here waits for the GPU except that collection."""
        pl, ctx = self.pl, self.ctx
        if not hasattr(self, "_id_buf"):
            self._id_buf = np.zeros(4 ** max(self.max_level - 2, 1), pl.TILE_ID_DTYPE)
        h2d = d2h = 0
        self.stats_checksum = 0.0      # sum of every (zmin, zmax) read back: must not depend on the partition
        pending = []
        for f, level, m0, n, s0, p0, pm0 in self.plan.batches(units, self.max_level):
            sc = self.scenes[f]
            ids = pl.make_tile_ids_range(level, m0, n, s0, p0, pm0, out=self._id_buf)
            ctx.pair_batch_ids(sc, self.elev, self.norm, ids)            # copies ids before it returns
            h2d += ids.nbytes
            if n >= 4096:     # the consumer's readback (TileSamplerZ): 8 bytes per tile, collected three
                if len(pending) == 3:                                    # read-backs later (ReadbackManager
                    st = ctx.elev_stats_readback_end(pending.pop(0))     # keeps several in flight)
                    d2h += st.nbytes
                    self.stats_checksum += float(st.astype(np.float64).sum())
                pending.append(ctx.elev_stats_readback_begin(self.elev, s0, n))
        for tk in pending:
            st = ctx.elev_stats_readback_end(tk)
            d2h += st.nbytes
            self.stats_checksum += float(st.astype(np.float64).sum())
        return h2d, d2h


# -------------------------------------------------------------- CPU baseline

def cpu_sample(max_level, face=1, nthreads=0, engine="port"):
    """One face of the same planet, levels 0..max_level, on the host cores (OpenMP over the tiles of a level).
    engine "port": the oracle (CPU restatement of the reference's GLSL path); "reference": the reference's OWN shader text
    compiled unchanged as C++ (oracle/_ref/libref_glsl.so; per-tile uniforms from the oracle: the reference's host code
    needs Ork) -- None when oracle/_ref is not there."""
    import orc
    orc.build()
    scene = orc.make_scene(W=101, gridMeshSize=24, rootQuadSize=PLANET_SIZE, face=face, flip=0,
                           noise_mode=1, no_clamp=0, noiseAmp=PLANET_AMP, sphere=1, elev_filter=1)
    t0 = time.perf_counter()
    r = orc.produce_quadtree(scene, max_level, nthreads) if engine == "port" else orc.glsl_produce_quadtree(scene, max_level, nthreads)
    if r is None:
        return None
    n, checksum, lo, hi = r
    dt = time.perf_counter() - t0
    return n, dt, (checksum, lo, hi)


def cpu_arm(nthreads, steps=1, warmup=0):
    """the CPU arm of the bench: the oracle port on one face, levels 0..7 per step -> (pairs, seconds, kind, sample text).
    The port is the FASTER of the two CPU implementations of the path this repo can run (see glsl_reference_sample): the
    conservative denominator for a speed-up."""
    level = 7
    for _ in range(min(warmup, 1)):
        cpu_sample(5, nthreads=nthreads)
    t_total, n_total = 0.0, 0
    for _ in range(max(steps, 1)):
        n, dt, _ = cpu_sample(level, nthreads=nthreads)
        t_total += dt
        n_total += n
    sample = "face 1 of the planet, levels 0..%d (%d pairs) per step, the oracle port, %d OpenMP threads" % (level, n_total // max(steps, 1), nthreads)
    return n_total, t_total, "port", sample


def glsl_reference_sample(nthreads, level=5):
    """the reference's OWN shader text (upsampleShader.glsl, normalShader.glsl compiled unchanged as C++ behind the shim:
    oracle/_ref/libref_glsl.so) on a smaller sample of the same workload -- reported beside the port, not used as the
    denominator: the shim's vector classes and texture emulation make it ~10 x slower than the port, which says more about the
    shim than about the reference.  None when oracle/_ref is not there."""
    r = cpu_sample(level, nthreads=nthreads, engine="reference")
    if r is None:
        return None
    n, dt, _ = r
    return {"value": n / dt, "unit": "pairs/s", "cores": nthreads, "kind": "reference",
            "sample": "face 1 of the planet, levels 0..%d (%d pairs), the reference's own upsampleShader.glsl / normalShader.glsl "
                      "compiled unchanged as C++ (oracle/_ref/libref_glsl.so), uniforms from the oracle, %d OpenMP threads" % (level, n, nthreads)}


def gpu_fingerprint(pl, ctx, max_level, face=1):
    """The same tiles as cpu_sample on the GPU (fused kernel, resident pool): the sum over all tiles of
    (zmin + zmax) and the global extremes of the per-tile statistics -- the oracle's checksum of checksums."""
    off = [sum(4 ** k for k in range(l)) for l in range(max_level + 2)]
    elev = ctx.pool(pl.POOL_ELEV, 101, off[max_level + 1])
    norm = ctx.pool(pl.POOL_NORM2, 97, off[max_level + 1])
    sc = pl.sweep_scene(noise_amp=PLANET_AMP, face=face, root_quad_size=PLANET_SIZE, sphere=1,
                        elev_filter=pl.FILTER_LINEAR, want_stats=1)
    for l in range(max_level + 1):
        ctx.produce_range(sc, elev, norm, l, 0, 4 ** l, off[l], off[l - 1] if l else 0, 0)
    st = ctx.elev_stats_range(elev, 0, off[max_level + 1]).astype(np.float64)
    out = float(st.sum()), float(st[:, 0].min()), float(st[:, 1].max())
    norm.close()
    elev.close()
    return out


ORTHO_BYTES = 196 * 196 * 4 + 100 * 100 * 4     # RGBA8 tile written + parent quadrant read


def ortho_lines(pl, ctx, torch, stream, peak, peak_kind, max_level=7, reps=3):
    """The ortho path (OrthoProducer) beside the headline: terrain3's hsv scene and a plain-noise
    scene, the full quadtree of levels 0..max_level of one face through pl_ortho_batch (host-built requests inside the
    timed region).  Auxiliary numbers: they do not enter `value`."""
    off = [(4 ** l - 1) // 3 for l in range(max_level + 2)]
    total = off[max_level + 1]
    ctx.ortho_noise_init(196)
    pool = ctx.pool(pl.POOL_ORTHO, 196, total)
    out = {}
    scenes = {"ortho_hsv": pl.ortho_scene(hsv=1, cnoise=(70, 80, 100), rnoise=(60, 150, 20), noise_amp=[255] * 17, face=1),
              "ortho_plain": pl.ortho_scene(hsv=0, cnoise=(127.5, 0, 0, 0), noise_amp=[0] + [255] * 16, face=3)}
    for name, sc in scenes.items():
        def sweep():
            for l in range(max_level + 1):
                ctx.ortho_batch(sc, pool, None, pl.ortho_make_requests_range(sc, l, 0, 4 ** l, out_slot0=off[l],
                                                                             parent_slot0=off[l - 1] if l else 0))
        sweep()
        ctx.sync()
        ctx.timing_collect()
        ctx.timing_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            sweep()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        k_ms, launches, tiles = ctx.timing_collect()["ortho"]
        ctx.timing_enable(False)
        gbs = ORTHO_BYTES * tiles / (k_ms * 1e-3) / 1e9
        out[name] = {"tiles_per_s": total / (ms * 1e-3), "tiles_per_sweep": total, "ms_per_sweep": ms,
                     "launches_per_sweep": launches // reps,
                     "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                                  "bytes_per_tile": ORTHO_BYTES, "peak_kind": peak_kind}}
    pool.close()
    return out


def host_cores():
    """Host threads the CPU arm may use: every core of the box (sched_getaffinity when it is narrower).  torchrun
    exports OMP_NUM_THREADS=1 to its workers; the arm passes this count to the oracle explicitly
    (omp_set_num_threads), which overrides the environment."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(args, rank):
    """--impl reference: the reference's algorithm for the path on the box's host cores, all of them (the oracle port, OpenMP
    over the tiles of a level).  One step = face 1 of the same planet, levels 0..7: 21 845 pairs, 98 % of them in levels
    5..7 (1 024 / 4 096 / 16 384 independent tiles per level: saturates any core count up to a few hundred).  The line also
    carries `reference_glsl`: the reference's own shader text compiled as C++, timed on a smaller sample."""
    if rank != 0:
        return
    cores = host_cores()
    n_total, t_total, kind, sample = cpu_arm(cores, args.steps, args.warmup)
    value = n_total / t_total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_total / max(args.steps, 1), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "demo-fractalplanet: 6 faces, levels 0..10, 8388606 pairs "
                                   "(bounded CPU sample per step)", "tile_w": 101, "normal_w": 97},
            "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": kind,
                             "sample": sample},
            "threads": cores, "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS"),
            "reference_glsl": glsl_reference_sample(cores),
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------- main

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--max-level", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-gather", action="store_true", help="N > 1: skip the `gather` record (tools/gather_tiles.py)")
    ap.add_argument("--no-configs", action="store_true", help="skip the sub-records of BASELINE configs 1, 3, 4, 5 (tools/configs.py)")
    ap.add_argument("--arith", default="fast", choices=["fast", "exact"],
                    help="arithmetic contract of the normal pass in the timed region (include/proland_b200.h: "
                         "PL_ARITH_FAST = within one unorm8 step of the canonical evaluation, PL_ARITH_EXACT = "
                         "bit-identical to it); elevations are bit-exact under both.  The other contract is "
                         "timed for one step and reported under `contracts`")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    # one rank per GPU on one box: the ranks share the host cores
    host_threads = max(1, (os.cpu_count() or 1) // max(world, 1))
    os.environ.setdefault("PL_HOST_THREADS", str(host_threads))
    import torch
    import torch.distributed as dist
    import proland_b200 as pl

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator is created: stdout carries the ONE
        # JSON line only, so the creation (init + a first collective) runs with fd 1 pointing at stderr
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            ctypes.CDLL(None).fflush(None)      # the banner sits in C stdio's buffer when stdout is a pipe
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = pl.Context(local_rank)
    stream = torch.cuda.Stream(device=local_rank)
    ctx.set_stream(stream.cuda_stream)       # torch events see this stream
    arith = pl.ARITH_FAST if args.arith == "fast" else pl.ARITH_EXACT
    sweep = PlanetSweep(pl, ctx, args.max_level, want_stats=1, arith=arith)
    my_units = sweep.plan.units_of_rank(sweep.units, rank, world)
    total_pairs = sweep.pairs_of(sweep.units, True)            # counted once per step, whole job
    launches0 = ctx.launches

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            sweep.run_device(my_units)
        barrier()
        ctx.timing_collect()
        ctx.timing_enable(True)
        clocks = ClockSampler(local_rank)
        if rank == 0:
            clocks.start()
        launches_before = ctx.launches
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record(stream)
        for _ in range(args.steps):
            sweep.run_device(my_units)
        ev1.record(stream)
        barrier()
        ms = ev0.elapsed_time(ev1)
        gpu_launches = ctx.launches - launches_before
        clock_rec = clocks.stop() if rank == 0 else None
        kt = ctx.timing_collect()
        ctx.timing_enable(False)

        # the other arithmetic contract, one step, same sweep (kernel events of the library)
        other = None
        other_name = "exact" if args.arith == "fast" else "fast"
        sweep.set_arith(pl.ARITH_EXACT if args.arith == "fast" else pl.ARITH_FAST)
        sweep.run_device(my_units[:1])
        barrier()
        ctx.timing_collect()
        ctx.timing_enable(True)
        ev4, ev5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev4.record(stream)
        sweep.run_device(my_units)
        ev5.record(stream)
        barrier()
        other = (ev4.elapsed_time(ev5), ctx.timing_collect())
        ctx.timing_enable(False)
        sweep.set_arith(arith)

        # e2e: host tile identities through the C ABI (pl_pair_batch_ids), stats read back; `steps` sweeps
        e2e = None
        if not args.no_e2e:
            sweep.run_host_ids(my_units[:1])
            barrier()
            ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            ev2.record(stream)
            h2d = d2h = 0
            for _ in range(args.steps):
                a, b = sweep.run_host_ids(my_units)
                h2d, d2h = h2d + a, d2h + b
            ev3.record(stream)
            barrier()
            e2e_s = max(time.perf_counter() - t0, 1e-3 * ev2.elapsed_time(ev3)) / args.steps
            e2e = (e2e_s, h2d // args.steps, d2h // args.steps)

    t = torch.tensor([ms, e2e[0] if e2e else 0.0, other[0]], dtype=torch.float64, device="cuda")
    c = torch.tensor([sweep.stats_checksum if e2e else 0.0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    ms_max, e2e_s_max, other_ms_max = float(t[0]), float(t[1]), float(t[2])

    # N > 1: gathering finished tiles and statistics (the one use the north star has for NCCL), and the same gather
    # without a collective -- the fused kernel's PUSH variant storing its normal tiles into the peers' pools
    gather = None
    if world > 1 and not args.no_gather:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import gather_tiles
        sweep.elev.close()      # the gather lays its own pools out for a whole quadtree
        sweep.norm.close()
        try:
            gather = gather_tiles.gather_record(ctx, torch, dist, stream, rank, world)
        except Exception as ex:         # a sub-record must not take the headline line with it
            gather = {"error": "%s: %s" % (type(ex).__name__, ex)}

    if rank == 0:
        peak, peak_kind = peaks()
        secs = ms_max * 1e-3
        value = total_pairs * args.steps / secs
        # dominant kernel = the one with the largest share of the step (the fused elevation+normal
        # kernel on this path; the separate passes appear when fusion is off)
        per_tile = {"pair": PAIR_BYTES, "elevation": ELEV_BYTES, "normal": NORM_BYTES}
        traffic = dict(TRAFFIC, pair=TRAFFIC["pair"] if args.arith == "fast" else TRAFFIC_EXACT_PAIR)
        roof = {}
        for k in per_tile:
            tot_ms, n_launch, n_tiles = kt[k]
            if n_launch == 0:
                continue
            gbs = per_tile[k] * n_tiles / (tot_ms * 1e-3) / 1e9 if tot_ms > 0 else 0.0
            per_launch = n_tiles / max(n_launch, 1)
            roof[k] = {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s",
                       "frac": gbs / peak,
                       # measured DRAM bytes (ncu --set full, profiles/) scaled to the average launch of this run
                       "traffic": traffic[k] * per_launch if traffic.get(k) else None,
                       "traffic_per_pair": traffic.get(k), "algorithmic_bytes_per_launch": per_tile[k] * per_launch,
                       "tiles_per_launch": per_launch, "launches": n_launch,
                       "avg_launch_ms": tot_ms / max(n_launch, 1), "share_of_step": tot_ms / ms_max,
                       "bytes_per_tile": per_tile[k], "peak_kind": peak_kind}
        dom = max(roof, key=lambda k: roof[k]["share_of_step"])
        line = {"metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": "demo-fractalplanet: 6 faces, levels 0..%d, %d pairs per step"
                                       % (args.max_level, total_pairs),
                           "tile_w": 101, "normal_w": 97, "partition": "96 level-2 subtrees, round-robin",
                           "arith": ARITH_NOTE[args.arith],
                           "l2": "each step writes > 1 TB, far larger than L2", "pool_slots": sweep.capacity},
                "hbm_gbs": value * PAIR_BYTES / 1e9, "hbm_frac": value * PAIR_BYTES / 1e9 / peak / world,
                "roofline": dict(roof[dom], kernel=dom), "kernels": roof,
                "gpu_launches": int(gpu_launches), "clocks": clock_rec}
        o_ms, o_launch, o_tiles = other[1]["pair"]
        o_gbs = PAIR_BYTES * o_tiles / (o_ms * 1e-3) / 1e9 if o_ms > 0 else 0.0
        line["contracts"] = {
            args.arith: {"value": value, "roofline_frac": roof[dom]["frac"], "steps": args.steps},
            other_name: {"value": total_pairs / (other_ms_max * 1e-3), "roofline_frac": o_gbs / peak, "steps": 1,
                         "arith": ARITH_NOTE[other_name]}}
        if e2e:
            line["e2e"] = {"value": total_pairs / e2e_s_max, "unit": "pairs/s",
                           "h2d_bytes_per_step": int(e2e[1]) * world, "d2h_bytes_per_step": int(e2e[2]) * world,
                           "path": "32-byte tile identities (HOST arrays) -> pl_pair_batch_ids (uniforms expanded on the "
                                   "device) -> fused kernel; per-tile (zmin, zmax) read back asynchronously",
                           "sweeps_timed": args.steps,
                           "stats_checksum": float(c[0]),
                           "stats_checksum_of": "sum of the (zmin, zmax) of every tile of levels 8..10 read back "
                                                "in the step: independent of the number of ranks"}
        if gather is not None:
            g = dict(gather)
            if "normals" in g:
                g["normals_GBps_per_rank"] = g["normals"]["GBps_per_rank"]
                g["identical_on_every_rank"] = bool(g.get("identical_to_single_gpu")) and \
                    bool(g.get("push_from_the_kernel", {}).get("identical_on_every_rank", False)) and \
                    bool(g.get("multicast_push_from_the_kernel", {"identical_on_every_rank": True})["identical_on_every_rank"])
            line["gather"] = g
        if world == 1 and not args.no_cpu_baseline:
            n, dt, (csum, clo, chi) = cpu_sample(7)
            line["cpu_baseline"] = {"value": n / dt, "unit": "pairs/s", "cores": host_cores(), "kind": "port",
                                    "sample": "face 1 of the same planet, levels 0..7 (%d pairs), oracle "
                                              "with OpenMP over the tiles of a level" % n,
                                    "reference_glsl": glsl_reference_sample(host_cores())}
            # the oracle run doubles as the checker: same tiles on the GPU, same checksum of checksums
            gsum, glo, ghi = gpu_fingerprint(pl, ctx, 7)
            line["parity"] = {"tiles": n, "vs": "oracle (cpu_baseline sample)",
                              "what": "sum over tiles of per-tile (zmin + zmax) of zm, and the global extremes",
                              "checksum_rel_err": abs(gsum - csum) / max(abs(csum), 1e-30),
                              "zmin_equal": glo == clo, "zmax_equal": ghi == chi,
                              "height_range_m": [clo, chi]}
        if world == 1 and not args.no_cpu_baseline:
            with torch.cuda.stream(stream):
                line["other_paths"] = ortho_lines(pl, ctx, torch, stream, peak, peak_kind)
        if world == 1 and not args.no_configs:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import configs
            with torch.cuda.stream(stream):
                line["configs"] = configs.run_all(pl, ctx, torch, stream, peak, peak_kind)
        print(json.dumps(line), flush=True)

    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
