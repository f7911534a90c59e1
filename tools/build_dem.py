"""The residual builder end to end: proland::preprocessSphericalDem of the host layer (map upload -> base grids
of the six cube faces in HBM -> per level: height tiles across the cube's edges, residuals, approximations -> DEM1..6.dat) against
the REFERENCE'S OWN builder (preprocess/terrain/*.cpp compiled unchanged: oracle/_ref/libref_hm.so, single-threaded as it is)
on the same synthetic equirectangular map.  The files must hold the same int16 tiles (compared tile by tile); prints one JSON
line with both times.

    python tools/build_dem.py [--tile 192] [--max-level 3] [--src 2048]
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "proland-4.0_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def tiles_of(orc, path):
    rd = orc.Resid(open(path, "rb").read())
    out = []
    for tid in range(rd.f.ntiles):
        raw, w, h = rd.inflate(tid)
        out.append(np.frombuffer(raw, np.int16))
    return (rd.f.minLevel, rd.f.maxLevel, rd.f.tileSize), out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--min-tile", type=int, default=24)
    ap.add_argument("--tile", type=int, default=192)
    ap.add_argument("--max-level", type=int, default=3)
    ap.add_argument("--src", type=int, default=2048)
    ap.add_argument("--no-reference", action="store_true")
    a = ap.parse_args()
    import orc
    import proland_host as ph
    sw, sh = a.src, a.src // 2
    rng = np.random.default_rng(42)
    yy, xx = np.mgrid[0:sh, 0:sw]
    src = (3000 * np.sin(xx * (2 * np.pi * 5 / sw)) * np.cos(yy * (np.pi * 7 / sh)) + 500 * np.sin(xx / 9.0) * np.sin(yy / 7.0)
           + 40 * rng.standard_normal((sh, sw))).astype(np.float32)
    base = a.tile << a.max_level
    rec = {"workload": "preprocessSphericalDem: %d x %d source map -> six cube faces of %d^2 base samples, tiles of %d (+5), "
                       "residual scale 1" % (sw, sh, base, a.tile)}
    with tempfile.TemporaryDirectory() as tmp:
        t = time.perf_counter()
        ph.preprocess_dem(src, a.min_tile, a.tile, a.max_level, os.path.join(tmp, "warm"), 1.0, True)     # context, kernels
        rec["device_first_call_s"] = time.perf_counter() - t
        t = time.perf_counter()
        ph.preprocess_dem(src, a.min_tile, a.tile, a.max_level, os.path.join(tmp, "dev"), 1.0, True)
        rec["device_s"] = time.perf_counter() - t
        hdr, ours = tiles_of(orc, os.path.join(tmp, "dev", "DEM3.dat"))
        rec["tiles_per_face"] = len(ours)
        rec["tiles"] = 6 * len(ours)
        rec["device_tiles_per_s"] = rec["tiles"] / rec["device_s"]
        if not a.no_reference and orc.hm() is not None:
            t = time.perf_counter()
            orc.ref_preprocess_dem(src, a.min_tile, a.tile, a.max_level, os.path.join(tmp, "ref"), os.path.join(tmp, "reftmp"), 1.0, True)
            rec["reference_s"] = time.perf_counter() - t
            rec["reference_kind"] = "the reference's own Preprocess.cpp / HeightMipmap.cpp (oracle/_ref/libref_hm.so), one thread, temporary tiles in memory"
            same = True
            for f in range(1, 7):
                h2, theirs = tiles_of(orc, os.path.join(tmp, "ref", "DEM%d.dat" % f))
                h1, mine = tiles_of(orc, os.path.join(tmp, "dev", "DEM%d.dat" % f))
                same = same and h1 == h2 and len(mine) == len(theirs) and all(np.array_equal(x, y) for x, y in zip(mine, theirs))
            rec["identical_to_the_reference"] = bool(same)
            rec["speedup"] = rec["reference_s"] / rec["device_s"]
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
