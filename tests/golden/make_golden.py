"""Regenerates tests/golden/*.json from the reference checkout (/root/reference) and the
reference's own noise.cpp compiled unchanged (oracle/_ref).  Run in the build container only;
the committed JSON files are what travels to the GPU box.

  python tests/golden/make_golden.py
"""
import base64
import hashlib
import json
import os
import struct
import sys
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference"


def dem_golden():
    path = os.path.join(REF, "src/terrain/examples/terrain4/DEM.dat")
    data = open(path, "rb").read()
    minL, maxL, ts, rootL, rtx, rty = struct.unpack_from("<6i", data, 0)
    scale, = struct.unpack_from("<f", data, 24)
    ntiles = minL + ((1 << (max(maxL - minL, 0) * 2 + 2)) - 1) // 3
    header = 28 + 8 * ntiles
    offs = struct.unpack_from("<%dI" % (2 * ntiles), data, 28)
    sha = []
    widths = []
    for t in range(ntiles):
        blob = data[header + offs[2 * t]: header + offs[2 * t + 1]]
        assert blob[:4] == b"II*\0"
        # single strip at byte 8, plain zlib stream (checked against the IFD in test_oracle.py)
        ifd, = struct.unpack_from("<I", blob, 4)
        n, = struct.unpack_from("<H", blob, ifd)
        tags = {}
        for i in range(n):
            tag, typ, cnt, val = struct.unpack_from("<HHII", blob, ifd + 2 + 12 * i)
            tags[tag] = val & 0xFFFF if typ == 3 else val
        raw = zlib.decompress(blob[tags[273]: tags[273] + tags[279]])
        assert len(raw) == tags[256] * tags[257] * 2
        sha.append(hashlib.sha1(raw).hexdigest())
        widths.append(tags[256])
    keep = [0, 1, 2, 3, 4, 5, 20, 100, 343]
    out = {
        "source": "src/terrain/examples/terrain4/DEM.dat",
        "md5": hashlib.md5(data).hexdigest(), "size": len(data),
        "header": {"minLevel": minL, "maxLevel": maxL, "tileSize": ts, "rootLevel": rootL,
                   "rootTx": rtx, "rootTy": rty, "scale": scale, "ntiles": ntiles, "header_bytes": header},
        "offsets_sha1": hashlib.sha1(data[28:header]).hexdigest(),
        "tile_sha1": sha, "tile_width": widths,
        "blobs": {str(t): base64.b64encode(data[header + offs[2 * t]: header + offs[2 * t + 1]]).decode()
                  for t in keep},
        "prefix": base64.b64encode(data[:header]).decode(),   # header + offset table (2780 bytes)
    }
    json.dump(out, open(os.path.join(HERE, "dem_dat.json"), "w"))
    print("dem_dat.json: %d tiles, %d blobs kept" % (ntiles, len(keep)))
    # the whole fixture, byte for byte (554 KB: header, offset table, 145 distinct TIFF/DEFLATE blobs): the GPU box has no
    # /root/reference, and the device decoders are checked on all 344 tiles (tests/test_gpu_parity.py)
    open(os.path.join(HERE, "terrain4_DEM.dat"), "wb").write(data)


def noise_golden():
    import ctypes as C
    import orc
    orc.build()
    R = orc.ref()
    assert R is not None, "oracle/_ref/libref_noise.so missing"
    seed = C.c_long(1234567)
    lcg = []
    fr = []
    for _ in range(8):
        s = C.c_long(seed.value)
        fr.append(R.ref_frandom(C.byref(s)))
        lcg.append(R.ref_lrandom(C.byref(seed)))
    pts = [(0.5, 1.0), (3.5, 7.0), (0.0, 0.5), (17.5, 1024.0), (4095.5, 2048.0), (-1023.5, 512.0),
           (1023.0, -1023.5), (100.5, 3071.0), (-4100.5, 12.0), (8191.5, -5000.0)]
    cn = [R.ref_cnoise2(x, y) for x, y in pts]
    json.dump({"source": "src/core/sources/proland/math/noise.{h,cpp} compiled unchanged (oracle/_ref)",
               "lcg_seed": 1234567, "lcg": lcg, "frandom_after_each": fr,
               "cnoise_points": pts, "cnoise": cn},
              open(os.path.join(HERE, "noise.json"), "w"))
    print("noise.json: lcg", lcg[:4], "cnoise", cn[:2])


if __name__ == "__main__":
    dem_golden()
    noise_golden()
