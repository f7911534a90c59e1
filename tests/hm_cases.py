"""Cases of the residual builder shared by the pin tests, the golden generator and the GPU tests.

A case = a synthetic source map + the arguments of preprocessDem / preprocessSphericalDem.  Three engines build it:
    reference   the reference's own Preprocess.cpp / HeightMipmap.cpp (oracle/_ref/libref_hm.so)     -> files
    oracle      the restatement oracle/orc_preprocess.c, tile by tile                                 -> tiles
    device      proland_host.preprocess_dem (pl_height_* + pl_residual_encode_batch + writer)         -> files
and every engine is reduced to the same record: per face, the header, the int16 tile of every tile id, and which ids
share a blob (the constant-tile rule).  TEST INFRASTRUCTURE."""
import hashlib
import os
import tempfile

import numpy as np

# name: (spherical, src_w, src_h, min_tile_size, tile_size, max_level, residual_scale, seed)
CASES = {
    "sphere_24_96_l1": (True, 256, 128, 24, 96, 1, 1.0, 5),
    "sphere_12_48_l2_scale2": (True, 256, 128, 12, 48, 2, 2.0, 6),
    "sphere_24_192_l1_flat_poles": (True, 512, 256, 24, 192, 1, 1.0, 7),
    "plane_24_96_l2": (False, 128, 128, 24, 96, 2, 1.0, 8),
    "plane_12_48_l1_scale4": (False, 64, 192, 12, 48, 1, 4.0, 9),
}


def source_map(name):
    spherical, sw, sh, _, _, _, _, seed = CASES[name]
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:sh, 0:sw]
    m = 2000 * np.sin(xx * (2 * np.pi * 3 / sw)) * np.cos(yy * (np.pi * 5 / sh)) + 300 * rng.standard_normal((sh, sw))
    if "flat_poles" in name:       # large all-zero regions: exercises the shared constant blob
        m[: sh // 3] = 0
        m[-sh // 4:] = 0
    return m.astype(np.float32)


def levels(name):
    _, _, _, minT, T, maxL, _, _ = CASES[name]
    B = T << maxL
    min_level = max_level = 0
    s = T
    while s > minT:
        min_level += 1
        s //= 2
    s = B
    while s > minT:
        max_level += 1
        s //= 2
    return B, min_level, max_level


def tile_list(name):
    """[(level, tx, ty, ts)] in tile id order"""
    _, _, _, minT, T, _, _, _ = CASES[name]
    B, _, max_level = levels(name)
    out = []
    for l in range(max_level + 1):
        nt = max(1, (B // T) >> (max_level - l))
        out += [(l, tx, ty, min(minT << l, T)) for ty in range(nt) for tx in range(nt)]
    return out


def record_of_files(oracle, name, folder):
    """{face: {"header": (minLevel, maxLevel, tileSize, scale), "tiles": [int16 arrays by id], "shared": [first id with the same blob]}}"""
    spherical = CASES[name][0]
    rec = {}
    for f in range(6 if spherical else 1):
        path = os.path.join(folder, "DEM%d.dat" % (f + 1) if spherical else "DEM.dat")
        rd = oracle.Resid(open(path, "rb").read())
        tiles, shared, first = [], [], {}
        offs = np.frombuffer(rd.buf[28:28 + 8 * rd.f.ntiles].tobytes(), np.uint32).reshape(-1, 2)
        for tid, (l, tx, ty, ts) in enumerate(tile_list(name)):
            assert rd.tile_id(l, tx, ty) == tid
            raw, w, h = rd.inflate(tid)
            shared.append(first.setdefault(tuple(offs[tid]), tid))
            # an all-zero tile shares the blob of the FIRST all-zero tile of the file, whatever that one's width
            # (HeightMipmap.cpp:601-611): only a tile with its own blob must have its own width
            assert w == h and (w == ts + 5 or shared[-1] != tid)
            tiles.append(np.frombuffer(raw, np.int16).reshape(w, w))
        rec[f] = {"header": (rd.f.minLevel, rd.f.maxLevel, rd.f.tileSize, float(rd.f.scale)), "tiles": tiles, "shared": shared}
    return rec


def reference_record(oracle, name):
    spherical, _, _, minT, T, maxL, scale, _ = CASES[name]
    with tempfile.TemporaryDirectory() as tmp:
        oracle.ref_preprocess_dem(source_map(name), minT, T, maxL, os.path.join(tmp, "dst"), os.path.join(tmp, "tmp"), scale, spherical)
        return record_of_files(oracle, name, os.path.join(tmp, "dst"))


def base_grids(oracle, name):
    spherical = CASES[name][0]
    B = levels(name)[0]
    src = source_map(name)
    return [oracle.spherical_base(src, f, B) for f in range(6)] if spherical else [oracle.plane_base(src, B)]


def oracle_record(oracle, name):
    spherical, _, _, minT, T, _, scale, _ = CASES[name]
    B, min_level, max_level = levels(name)
    faces = base_grids(oracle, name)
    rec = {}
    for f in range(len(faces)):
        approx, tiles, zero = {}, [], None
        shared = []
        for tid, (l, tx, ty, ts) in enumerate(tile_list(name)):
            tile = oracle.hm_get_tile(faces, max_level, minT, T, l, f, tx, ty, scale)
            if l == 0:      # produceTile: short(roundf(h / scale)), half away from zero; its own approximation
                t = tile[:ts + 5, :ts + 5]
                r = (np.sign(t) * np.floor(np.abs(t) + np.float32(0.5))).astype(np.int16)
                approx[(0, 0, 0)] = tile
            else:
                r, a, _, _ = oracle.hm_encode_tile(approx[(l - 1, tx // 2, ty // 2)], tile, ts, tx, ty)
                approx[(l, tx, ty)] = a
            if not r.any():
                zero = tid if zero is None else zero
                shared.append(zero)
                tiles.append(r if zero == tid else tiles[zero])
            else:
                shared.append(tid)
                tiles.append(r)
        rec[f] = {"header": (min_level, max_level, T, float(scale)), "tiles": tiles, "shared": shared}
    return rec


def digest(rec):
    """one sha1 per face over header, sharing structure and tile bytes"""
    out = {}
    for f in sorted(rec):
        h = hashlib.sha1()
        h.update(repr(tuple(rec[f]["header"])).encode())
        h.update(np.asarray(rec[f]["shared"], np.int32).tobytes())
        for t in rec[f]["tiles"]:
            h.update(np.ascontiguousarray(t, np.int16).tobytes())
        out[str(f)] = h.hexdigest()[:20]
    return out


def first_difference(a, b):
    for f in sorted(a):
        if tuple(a[f]["header"]) != tuple(b[f]["header"]):
            return "face %d: header %r != %r" % (f, a[f]["header"], b[f]["header"])
        for tid, (x, y) in enumerate(zip(a[f]["tiles"], b[f]["tiles"])):
            if x.shape != y.shape or not np.array_equal(x, y):
                d = np.argwhere(x != y) if x.shape == y.shape else []
                return "face %d tile %d: %d samples differ, first %s" % (f, tid, len(d), d[:3].tolist() if len(d) else "shape")
        if list(a[f]["shared"]) != list(b[f]["shared"]):
            return "face %d: blob sharing differs" % f
    return None
