"""Config 5 (BASELINE.json): asynchronous camera fly-through, demo-earth-srtm-async style.

A deterministic camera path over one cube face of the synthetic SRTM-shaped planet (config 3:
residual files in the reference's container format, delta = 2, flip, NEAREST elevation storage,
sphere-deformed normals).  Every frame runs what the reference runs per frame:

    TerrainNode::update  -> TerrainQuad split rule (TerrainQuad.cpp:81-173)
    TileSampler::update  -> putTiles / getTiles / prefetch against the TileCache LRU pools
                            (nTiles = 1296, earth-srtm.xml; prefetchRate 2 / prefetchQueue 64,
                            earth-srtm-async.xml:27)
    Scheduler::run       -> BatchScheduler: one kernel launch per producer per dependency wave

and reports tiles/s, cache miss rate and the p50 / p99 frame production time (host clock around
the frame, device synchronised).

    python tools/flythrough.py [--frames 300] [--async] [--ntiles 1296]
"""
import argparse
import json
import math
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "proland-4.0_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

R = 6360000.0
NOISE = "0,0,0,0,0,0,0,0,0,0,0,5,2.5,1,0.5,0.25,0.1,0.05,0.025,0.01,0.01,0.005,0.005"


def archive(ntiles=1296, asynchronous=False, face=2):
    sched = 'prefetchRate="2" prefetchQueue="64"' if asynchronous else ""
    return """<?xml version="1.0" ?>
<archive>
    <multithreadScheduler name="defaultScheduler" nthreads="3" fps="0" %s/>
    <tileCache name="groundResiduals" scheduler="defaultScheduler">
        <cpuFloatTileStorage tileSize="197" channels="1" capacity="%d"/>
    </tileCache>
    <residualProducer name="groundResiduals%d" cache="groundResiduals" file="dem/DEM%d.dat" delta="2"/>
    <tileCache name="groundElevations" scheduler="defaultScheduler">
        <gpuTileStorage tileSize="101" nTiles="%d"
            internalformat="RGB32F" format="RGB" type="FLOAT" min="NEAREST" mag="NEAREST"/>
    </tileCache>
    <elevationProducer name="groundElevations%d" cache="groundElevations" residuals="groundResiduals%d" flip="true"
        noise="%s"/>
    <tileCache name="groundNormals" scheduler="defaultScheduler">
        <gpuTileStorage tileSize="97" nTiles="%d"
            internalformat="RG8" format="RG" type="FLOAT" min="LINEAR" mag="LINEAR"/>
    </tileCache>
    <normalProducer name="groundNormals%d" cache="groundNormals" elevations="groundElevations%d" deform="sphere"/>
</archive>""" % (sched, max(ntiles // 2, 64), face, face, ntiles, face, face, NOISE, ntiles, face, face)


def write_residuals(data_dir, face=2, max_level=6, seed=20240612):
    """a synthetic DEM<face>.dat in the reference's container format (tests/resid_synth.py)"""
    import resid_synth as rs
    os.makedirs(os.path.join(data_dir, "dem"), exist_ok=True)
    data, _ = rs.container(min_level=3, max_level=max_level, tile_size=192, scale=1.0, seed=seed + face)
    path = os.path.join(data_dir, "dem", "DEM%d.dat" % face)
    with open(path, "wb") as f:
        f.write(data)
    return data


def camera_path(frames):
    """a low pass over the face: a Lissajous curve in the local plane, altitude dipping from 400 km to 2 km"""
    for k in range(frames):
        u = k / max(frames - 1, 1)
        x = 0.55 * R * math.sin(2.0 * math.pi * (0.7 * u + 0.05))
        y = 0.55 * R * math.sin(2.0 * math.pi * (1.1 * u))
        z = 2000.0 + 398000.0 * (0.5 + 0.5 * math.cos(2.0 * math.pi * 1.5 * u)) ** 2
        yield x, y, z


def run(frames=300, asynchronous=False, ntiles=1296, max_level=12, data_dir=None, face=2, on_frame=None):
    import numpy as np
    import proland_host as ph
    tmp = None
    if data_dir is None:
        tmp = tempfile.TemporaryDirectory()
        data_dir = tmp.name
        write_residuals(data_dir, face)
    scene = ph.Scene(archive(ntiles, asynchronous, face), data_dir=data_dir)
    normals = scene.producer("groundNormals%d" % face)
    elevations = scene.producer("groundElevations%d" % face)
    residuals = scene.producer("groundResiduals%d" % face)
    sched = scene.scheduler("defaultScheduler")
    terrain = ph.Terrain(R, zmin=0.0, zmax=10000.0, split_factor=2.0, max_level=max_level)
    samplers = [ph.Sampler("elevationSampler", elevations, asynchronous), ph.Sampler("fragmentNormalSampler", normals, asynchronous)]
    split = ph.lib().plh_split_distance(2.0, 1024.0, math.radians(80.0))
    times, quads, made = [], [], []
    parts = [0.0, 0.0, 0.0]     # quadtree update, samplers + scheduler, device sync
    launches0 = ph.lib().plh_device_launches(-1)
    for k, cam in enumerate(camera_path(frames)):
        before = elevations.counts()[0] + normals.counts()[0] + residuals.counts()[0]
        t0 = time.perf_counter()
        nq = terrain.update(*cam, split_dist=split)
        t1 = time.perf_counter()
        ph.frame_update(sched, terrain, samplers)
        t2 = time.perf_counter()
        ph.lib().plh_device_sync(-1)
        t3 = time.perf_counter()
        times.append(t3 - t0)
        parts[0] += t1 - t0
        parts[1] += t2 - t1
        parts[2] += t3 - t2
        quads.append(nq)
        made.append(elevations.counts()[0] + normals.counts()[0] + residuals.counts()[0] - before)
        if on_frame is not None:
            on_frame(k, cam, terrain, elevations, normals)
    stats = {n: scene.cache(n).stats() for n in ("groundElevations", "groundNormals", "groundResiduals")}
    total = sum(made)
    t = np.array(times)
    busy = t[np.array(made) > 0]
    result = {
        "workload": "config 5 fly-through: face %d, %d frames, LRU pools of %d tiles, %s" % (face, frames, ntiles, "async (prefetchRate 2, prefetchQueue 64)" if asynchronous else "synchronous"),
        "frames": frames, "tiles_made": int(total), "tiles_per_s": total / float(t.sum()),
        "tiles_per_s_busy_frames": float(sum(made) / busy.sum()) if busy.size else 0.0,
        "frame_ms_p50": float(np.percentile(t, 50) * 1e3), "frame_ms_p99": float(np.percentile(t, 99) * 1e3),
        "frame_ms_max": float(t.max() * 1e3), "quads_mean": float(np.mean(quads)), "quads_max": int(max(quads)),
        "tiles_per_frame_max": int(max(made)),
        "miss_rate": {n: s["misses"] / max(s["queries"], 1) for n, s in stats.items()},
        "host_seconds": {"terrain_update": parts[0], "samplers_and_scheduler": parts[1], "device_sync": parts[2]},
        "kernel_launches": int(ph.lib().plh_device_launches(-1) - launches0),
        "tiles_per_launch": total / max(int(ph.lib().plh_device_launches(-1) - launches0), 1),
    }
    for s in samplers:
        s.close()
    terrain.close()
    scene.close()
    if tmp is not None:
        tmp.cleanup()
    return result


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=300)
    ap.add_argument("--async", dest="asynchronous", action="store_true")
    ap.add_argument("--ntiles", type=int, default=1296)
    ap.add_argument("--max-level", type=int, default=12)
    a = ap.parse_args()
    print(json.dumps(run(a.frames, a.asynchronous, a.ntiles, a.max_level)))
