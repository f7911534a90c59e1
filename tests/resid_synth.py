"""Synthetic residual containers in the exact format ResidualProducer reads
(src/terrain/doc/overview.txt:147-216; writer preprocess/terrain/HeightMipmap.cpp:561-655):

  header  : 6 int32 (minLevel, maxLevel, tileSize, rootLevel, rootTx, rootTy) + float32 scale
  offsets : ntiles x (begin, end) uint32, relative to the end of the offset table
  blobs   : one little-endian TIFF per tile: 1 strip, 2 x 8-bit samples (= LE int16),
            MINISBLACK, compression 32946 (DEFLATE), strip at byte 8, IFD after the strip

TEST INFRASTRUCTURE (also used by bench-style residual workloads): real DEM1..6.dat are
download-only (README.TXT:50-57), so config 3 runs on files made here.
"""
import struct
import zlib

import numpy as np


def hand_made_zlib_stream(complete):
    """a zlib stream of ONE dynamic block holding the two literals 'A' 'A' (one int16 texel 0x4141), its bits written by
    hand.  complete=False: the literal/length code has only the lengths {65: 2, 256: 2} -- an INCOMPLETE code, which zlib
    (and therefore the reference's libtiff) rejects ("invalid literal/lengths set"); complete=True adds symbol 66 with
    length 1, which makes the same stream valid."""
    bits = []

    def put(value, n):               # n bits, least significant first (header fields, extra bits)
        bits.extend((value >> k) & 1 for k in range(n))

    def code(value, n):              # a Huffman code, most significant bit first
        bits.extend((value >> (n - 1 - k)) & 1 for k in range(n))

    put(1, 1)                        # BFINAL
    put(2, 2)                        # BTYPE = dynamic
    put(0, 5)                        # HLIT: 257 literal/length codes
    put(0, 5)                        # HDIST: 1 distance code
    # code-length code: symbols 18 (1 bit), 0, 1 or 2 (2 bits each): complete.  HCLEN covers the order up to that symbol
    order = [16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15]
    third = 2 if not complete else 1
    cl = {18: 1, 0: 2, third: 2}
    ncl = 16 if not complete else 18
    put(ncl - 4, 4)
    for sym in order[:ncl]:
        put(cl.get(sym, 0), 3)
    # canonical code-length codes: 18 -> 0; then by symbol value among the 2-bit ones
    two = sorted(k for k in cl if cl[k] == 2)
    cc = {18: (0, 1), two[0]: (2, 2), two[1]: (3, 2)}

    def zeros(n):                    # n zeros, 11 <= n <= 138, by symbol 18
        code(*cc[18])
        put(n - 11, 7)

    if not complete:                 # lengths: 65 zeros, 2, 190 zeros, 2 | distance: 0
        zeros(65); code(*cc[2]); zeros(138); zeros(52); code(*cc[2]); code(*cc[0])
        lit = {65: (0, 2), 256: (1, 2)}
    else:                            # lengths {65: 1, 66: 0 ..., 256: 1}: a complete 1-bit code
        zeros(65); code(*cc[1]); zeros(138); zeros(52); code(*cc[1]); code(*cc[0])
        lit = {65: (0, 1), 256: (1, 1)}
    code(*lit[65]); code(*lit[65]); code(*lit[256])
    while len(bits) % 8:
        bits.append(0)
    body = bytes(sum(b << k for k, b in enumerate(bits[i:i + 8])) for i in range(0, len(bits), 8))
    return b"\x78\x9c" + body + struct.pack(">I", zlib.adler32(b"AA"))


def tiff_blob(tile, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, compression=32946, strip=None):
    """tile: (w, w) int16 -> TIFF bytes (strip: a ready-made compressed strip to wrap instead)"""
    tile = np.ascontiguousarray(tile, "<i2")
    w = tile.shape[0]
    raw = tile.tobytes()
    if strip is not None:
        pass
    elif compression == 1:
        strip = raw
    else:
        co = zlib.compressobj(level, zlib.DEFLATED, 15, 8, strategy)
        strip = co.compress(raw) + co.flush()
    ifd_off = 8 + len(strip)
    if ifd_off % 2:
        strip += b"\0"
        ifd_off += 1
    tags = [(256, 4, 1, w), (257, 4, 1, w), (258, 3, 2, 8 | (8 << 16)), (259, 3, 1, compression),
            (262, 3, 1, 1), (273, 4, 1, 8), (274, 3, 1, 4), (277, 3, 1, 2), (279, 4, 1, len(strip)),
            (284, 3, 1, 1)]
    out = b"II*\0" + struct.pack("<I", ifd_off) + strip + struct.pack("<H", len(tags))
    for tag, typ, cnt, val in tags:
        out += struct.pack("<HHII", tag, typ, cnt, val)
    return out + struct.pack("<I", 0)


def tile_width(min_level, tile_size, l):
    return (tile_size >> (min_level - l) if l < min_level else tile_size) + 5


def n_tiles(min_level, max_level):
    return min_level + ((1 << (max(max_level - min_level, 0) * 2 + 2)) - 1) // 3


def tile_id(min_level, l, tx, ty):
    if l < min_level:
        return l
    d = l - min_level
    return min_level + tx + ty * (1 << d) + ((1 << (2 * d)) - 1) // 3


def fractal_tile(rng, w, amp):
    """smooth-ish int16 residuals: a few octaves of bilinear noise, amplitude ~amp"""
    z = np.zeros((w, w), np.float64)
    step, a = max(w // 2, 1), float(amp)
    while step >= 1 and a >= 0.5:
        n = w // step + 2
        coarse = rng.normal(0, a, (n, n))
        yy, xx = np.mgrid[0:w, 0:w] / step
        x0, y0 = xx.astype(int), yy.astype(int)
        fx, fy = xx - x0, yy - y0
        z += ((1 - fx) * (1 - fy) * coarse[y0, x0] + fx * (1 - fy) * coarse[y0, x0 + 1]
              + (1 - fx) * fy * coarse[y0 + 1, x0] + fx * fy * coarse[y0 + 1, x0 + 1])
        step //= 2
        a *= 0.55
    return np.clip(np.rint(z), -32768, 32767).astype(np.int16)


def container(min_level=3, max_level=5, tile_size=192, root=(0, 0, 0), scale=1.0, seed=20240612,
              amps=(359, 43, 15, 3.9, 1.0, 0.5, 0.25, 0.12), zero_fraction=0.15, level=6):
    """-> (file bytes, {tile id: int16 array})"""
    rng = np.random.default_rng(seed)
    nt = n_tiles(min_level, max_level)
    tiles, blobs = {}, {}
    zero_blob = None
    for l in range(0, max_level + 1):
        cnt = 1 if l < min_level else 1 << (l - min_level)
        w = tile_width(min_level, tile_size, l)
        for ty in range(cnt):
            for tx in range(cnt):
                tid = tile_id(min_level, l, tx, ty)
                if l >= min_level and rng.random() < zero_fraction:
                    t = np.zeros((w, w), np.int16)          # constant tiles share one blob
                    if zero_blob is None:
                        zero_blob = tiff_blob(t, level)
                    blobs[tid] = zero_blob
                else:
                    t = fractal_tile(rng, w, amps[min(l, len(amps) - 1)])
                    blobs[tid] = tiff_blob(t, level)
                tiles[tid] = t
    body = b""
    offsets = []
    pos_of = {}
    for tid in range(nt):
        b = blobs[tid]
        if id(b) in pos_of:                                   # de-duplicated (HeightMipmap.cpp:601-611)
            offsets += list(pos_of[id(b)])
            continue
        pos_of[id(b)] = (len(body), len(body) + len(b))
        offsets += [len(body), len(body) + len(b)]
        body += b
    head = struct.pack("<6if", min_level, max_level, tile_size, root[0], root[1], root[2], scale)
    return head + struct.pack("<%dI" % len(offsets), *offsets) + body, tiles


def container_from_tiles(tiles, min_level, max_level, tile_size, root=(0, 0, 0), scale=1.0, level=6):
    """{tile id: (w, w) int16} -> file bytes, laid out like HeightMipmap::generate (offset table in id
    order, all-zero tiles sharing the first all-zero blob, HeightMipmap.cpp:601-611).  The blob order in
    the body is id order here (the reference writes Lebesgue order; readers only use the offset table)."""
    nt = n_tiles(min_level, max_level)
    body, offsets, zero = b"", [], None
    for tid in range(nt):
        t = np.ascontiguousarray(tiles[tid], np.int16)
        if not t.any():
            if zero is None:
                b = tiff_blob(t, level)
                zero = (len(body), len(body) + len(b))
                body += b
            offsets += list(zero)
            continue
        b = tiff_blob(t, level)
        offsets += [len(body), len(body) + len(b)]
        body += b
    head = struct.pack("<6if", min_level, max_level, tile_size, root[0], root[1], root[2], scale)
    return head + struct.pack("<%dI" % len(offsets), *offsets) + body


# ------------------------------------------------------------------------------------- ortho residual files

def ortho_tiff_blob(tile, level=6, compression=32946):
    """tile: (w, w, channels) uint8 -> the TIFF ColorMipmap::produceTile writes
    (preprocess/terrain/ColorMipmap.cpp:312-325): channels x 8-bit samples, contiguous, one DEFLATE strip.
    With more than two samples the BitsPerSample values do not fit the tag's value field and sit at an offset,
    as libtiff writes them."""
    tile = np.ascontiguousarray(tile, np.uint8)
    w, ch = tile.shape[0], tile.shape[2]
    raw = tile.tobytes()
    strip = raw if compression == 1 else zlib.compress(raw, level)
    strip_len = len(strip)
    if len(strip) % 2:
        strip += b"\0"
    extra = b""
    bps_val = 8 | (8 << 16) if ch == 2 else 8
    extra_off = 8 + len(strip)
    if ch > 2:
        extra = struct.pack("<%dH" % ch, *([8] * ch))
        bps_val = extra_off
    ifd_off = extra_off + len(extra)
    tags = [(256, 4, 1, w), (257, 4, 1, w), (258, 3, ch, bps_val), (259, 3, 1, compression),
            (262, 3, 1, 1 if ch == 1 else 2), (273, 4, 1, 8), (274, 3, 1, 4), (277, 3, 1, ch),
            (279, 4, 1, strip_len), (284, 3, 1, 1)]
    out = b"II*\0" + struct.pack("<I", ifd_off) + strip + extra + struct.pack("<H", len(tags))
    for tag, typ, cnt, val in tags:
        out += struct.pack("<HHII", tag, typ, cnt, val)
    return out + struct.pack("<I", 0)


def ortho_container(tiles, max_level, tile_size=192, channels=3, root=(0, 0, 0), flags=0, level=6):
    """tiles: {(level, tx, ty): (w, w, channels) uint8} for every tile of levels 0..max_level -> file bytes in the
    format OrthoCPUProducer reads (OrthoCPUProducer.cpp:84-118): 7 int32, (begin, end) int64 per tile id, blobs."""
    ntiles = (4 ** (max_level + 1) - 1) // 3
    blobs, offs = [], []
    pos = 0
    for tid in range(ntiles):
        l = 0
        while (4 ** (l + 1) - 1) // 3 <= tid:
            l += 1
        k = tid - (4 ** l - 1) // 3
        tx, ty = k % (1 << l), k // (1 << l)
        b = ortho_tiff_blob(tiles[(l, tx, ty)], level)
        blobs.append(b)
        offs.append((pos, pos + len(b)))
        pos += len(b)
    head = struct.pack("<7i", max_level, tile_size, channels, root[0], root[1], root[2], flags)
    table = b"".join(struct.pack("<qq", a, b) for a, b in offs)
    return head + table + b"".join(blobs)


def ortho_container_blob(file_bytes, level, tx, ty):
    """what the host side does before the decode call: header, offset table, tile id -> the blob's bytes"""
    max_level = struct.unpack_from("<i", file_bytes, 0)[0]
    ntiles = (4 ** (max_level + 1) - 1) // 3
    header = 28 + 16 * ntiles
    tid = tx + ty * (1 << level) + (4 ** level - 1) // 3
    a, b = struct.unpack_from("<qq", file_bytes, 28 + 16 * tid)
    return file_bytes[header + a:header + b]
