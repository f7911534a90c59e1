/*
 * orc_residual.c -- ORACLE (test infrastructure, never shipped): the
 * ResidualProducer pass (residual container, tile decode, root composition).
 *
 * Restates:
 *   terrain/sources/proland/dem/ResidualProducer.cpp:70-129  (header, offsets)
 *   terrain/sources/proland/dem/ResidualProducer.cpp:161-232 (hasTile, doCreateTile)
 *   terrain/sources/proland/dem/ResidualProducer.cpp:253-340 (tile id/size, readTile)
 *   terrain/sources/proland/dem/ResidualProducer.cpp:342-384 (upsample)
 *   file format: src/terrain/doc/overview.txt:170-216 and the writer
 *   preprocess/terrain/HeightMipmap.cpp:561-655.
 *
 * Third-party arithmetic absent from the reference tree: libtiff 3.x + zlib 1.x
 * (TIFFReadEncodedStrip over a DEFLATE strip).  The reference pins no version
 * (no lock file; only bin/libtiff3.dll, bin/zlib1.dll).  DEFLATE is lossless
 * (RFC 1950/1951), so the system zlib used here yields the same bytes; pinned
 * by the sha1 KATs of src/terrain/examples/terrain4/DEM.dat (tests/golden).
 */
#include "orc.h"
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

static uint32_t rd32(const uint8_t *p) { return (uint32_t) p[0] | ((uint32_t) p[1] << 8) | ((uint32_t) p[2] << 16) | ((uint32_t) p[3] << 24); }
static uint16_t rd16(const uint8_t *p) { return (uint16_t) (p[0] | (p[1] << 8)); }
static int imax(int a, int b) { return a > b ? a : b; }

int orc_resid_open(const uint8_t *data, size_t size, int deltaLevel, float zscale,
                   orc_resid_file *f)
{
    memset(f, 0, sizeof(*f));
    if (size < 28) return -1;
    f->minLevel = (int) rd32(data + 0);
    f->maxLevel = (int) rd32(data + 4);
    f->tileSize = (int) rd32(data + 8);
    f->rootLevel = (int) rd32(data + 12);
    f->rootTx = (int) rd32(data + 16);
    f->rootTy = (int) rd32(data + 20);
    memcpy(&f->scale, data + 24, 4);
    f->deltaLevel = f->rootLevel == 0 ? deltaLevel : 0;
    f->scale = f->scale * zscale;
    f->ntiles = f->minLevel + ((1 << (imax(f->maxLevel - f->minLevel, 0) * 2 + 2)) - 1) / 3;
    f->header = (uint32_t) (4 + 4 * (6 + f->ntiles * 2));
    if (size < f->header) return -2;
    f->offsets = (const uint32_t *) (data + 28);
    f->data = data;
    f->size = size;
    return 0;
}

int orc_resid_has_tile(const orc_resid_file *f, int level, int tx, int ty)
{
    int l = level + f->deltaLevel - f->rootLevel;
    if (l >= 0 && (tx >> l) == f->rootTx && (ty >> l) == f->rootTy) {
        if (l <= f->maxLevel) return 1;
        for (int i = 0; i < f->nchildren; ++i) {
            if (orc_resid_has_tile(f->children[i], level + f->deltaLevel, tx, ty)) return 1;
        }
    }
    return 0;
}

int orc_resid_tile_size(const orc_resid_file *f, int l)
{
    return l < f->minLevel ? f->tileSize >> (f->minLevel - l) : f->tileSize;
}

int orc_resid_tile_id(const orc_resid_file *f, int l, int tx, int ty)
{
    if (l < f->minLevel) return l;
    int d = imax(l - f->minLevel, 0);
    return f->minLevel + tx + ty * (1 << d) + ((1 << (2 * d)) - 1) / 3;
}

const uint8_t *orc_resid_blob(const orc_resid_file *f, int tileid, uint32_t *size)
{
    uint32_t a, b;
    memcpy(&a, &f->offsets[2 * tileid], 4);
    memcpy(&b, &f->offsets[2 * tileid + 1], 4);
    *size = b - a;
    return f->data + f->header + a;
}

/* Minimal baseline-TIFF reader: little- or big-endian header, first IFD, one
 * strip (tags 273/279), compression 8 or 32946 (both are zlib streams), no
 * predictor.  Returns the strip exactly as TIFFReadEncodedStrip would. */
long orc_tiff_inflate(const uint8_t *blob, uint32_t size, uint8_t *raw, size_t cap,
                      int *width, int *height)
{
    if (size < 8 || blob[0] != 'I' || blob[1] != 'I' || rd16(blob + 2) != 42) return -1;
    uint32_t ifd = rd32(blob + 4);
    if (ifd + 2 > size) return -2;
    int n = rd16(blob + ifd);
    uint32_t strip_off = 0, strip_len = 0, comp = 1, w = 0, h = 0, predictor = 1;
    for (int i = 0; i < n; ++i) {
        const uint8_t *e = blob + ifd + 2 + 12 * i;
        if ((size_t) (e - blob) + 12 > size) return -3;
        uint16_t tag = rd16(e), type = rd16(e + 2);
        uint32_t val = (type == 3) ? rd16(e + 8) : rd32(e + 8);
        switch (tag) {
        case 256: w = val; break;
        case 257: h = val; break;
        case 259: comp = val; break;
        case 273: strip_off = val; break;
        case 279: strip_len = val; break;
        case 317: predictor = val; break;
        default: break;
        }
    }
    if (width) *width = (int) w;
    if (height) *height = (int) h;
    if (predictor != 1) return -4;
    if (strip_off + strip_len > size) return -5;
    if (comp == 1) {
        if (strip_len > cap) return -6;
        memcpy(raw, blob + strip_off, strip_len);
        return (long) strip_len;
    }
    if (comp != 8 && comp != 32946) return -7;
    uLongf out = (uLongf) cap;
    int rc = uncompress(raw, &out, blob + strip_off, strip_len);
    if (rc != Z_OK) return -100 + rc;
    return (long) out;
}

int orc_resid_read_tile(const orc_resid_file *f, int l, int tx, int ty,
                        const float *tile, float *result)
{
    const int n = f->tileSize + 5;
    const int w = orc_resid_tile_size(f, l) + 5;
    const int id = orc_resid_tile_id(f, l, tx, ty);
    uint32_t bsize;
    const uint8_t *blob = orc_resid_blob(f, id, &bsize);
    size_t cap = (size_t) n * n * 2;
    uint8_t *raw = (uint8_t *) malloc(cap);
    long got = orc_tiff_inflate(blob, bsize, raw, cap, NULL, NULL);
    if (got != (long) w * w * 2) { free(raw); return -1; }
    for (int j = 0; j < w; ++j) {
        for (int i = 0; i < w; ++i) {
            int off = 2 * (i + j * w);
            int toff = i + j * n;
            short z = (short) ((short) raw[off + 1] << 8 | (short) raw[off]);
            float zs = (float) z * f->scale;
            result[toff] = tile ? tile[toff] + zs : zs;
        }
    }
    free(raw);
    return 0;
}

void orc_resid_upsample(const orc_resid_file *f, int l, int tx, int ty,
                        const float *parentTile, float *result)
{
    const int n = f->tileSize + 5;
    const int ts = orc_resid_tile_size(f, l);
    const int px = 1 + (tx % 2) * ts / 2;
    const int py = 1 + (ty % 2) * ts / 2;
#define P(a, b) parentTile[(a) + (b) * n]
    for (int j = 0; j <= ts + 4; ++j) {
        for (int i = 0; i <= ts + 4; ++i) {
            const int cx = i / 2 + px, cy = j / 2 + py;
            float z;
            if (j % 2 == 0) {
                if (i % 2 == 0) {
                    z = P(cx, cy);
                } else {
                    float z0 = P(cx - 1, cy), z1 = P(cx, cy), z2 = P(cx + 1, cy), z3 = P(cx + 2, cy);
                    z = ((z1 + z2) * 9 - (z0 + z3)) / 16;
                }
            } else {
                if (i % 2 == 0) {
                    float z0 = P(cx, cy - 1), z1 = P(cx, cy), z2 = P(cx, cy + 1), z3 = P(cx, cy + 2);
                    z = ((z1 + z2) * 9 - (z0 + z3)) / 16;
                } else {
                    z = 0.0f;
                    for (int dj = -1; dj <= 2; ++dj) {
                        float fw = (dj == -1 || dj == 2) ? -1 / 16.0f : 9 / 16.0f;
                        for (int di = -1; di <= 2; ++di) {
                            float gw = (di == -1 || di == 2) ? -1 / 16.0f : 9 / 16.0f;
                            z += fw * gw * P(cx + di, cy + dj);
                        }
                    }
                }
            }
            result[i + j * n] = z;
        }
    }
#undef P
}

int orc_resid_create_tile(const orc_resid_file *f, int level, int tx, int ty, float *out)
{
    int l = level + f->deltaLevel - f->rootLevel;
    if (l >= 0 && (tx >> l) == f->rootTx && (ty >> l) == f->rootTy) {
        if (l > f->maxLevel) {
            for (int i = 0; i < f->nchildren; ++i) {
                orc_resid_create_tile(f->children[i], level + f->deltaLevel, tx, ty, out);
            }
            return 0;
        }
    } else {
        return 0;
    }
    tx -= f->rootTx << l;
    ty -= f->rootTy << l;
    if (f->deltaLevel > 0 && l == f->deltaLevel) {
        const int n = f->tileSize + 5;
        float *tmp = (float *) malloc(sizeof(float) * n * n);
        int rc = orc_resid_read_tile(f, 0, 0, 0, NULL, out);
        for (int i = 1; i <= f->deltaLevel && rc == 0; ++i) {
            orc_resid_upsample(f, i, 0, 0, out, tmp);
            rc = orc_resid_read_tile(f, i, 0, 0, tmp, out);
        }
        free(tmp);
        return rc;
    }
    return orc_resid_read_tile(f, l, tx, ty, NULL, out);
}
