/*
 * orc_preprocess.c -- TEST INFRASTRUCTURE: CPU restatement of the residual-pyramid builder, the step
 * BEFORE the hot path.  Reference:
 *
 *   terrain/sources/proland/preprocess/terrain/HeightMipmap.cpp
 *     :449-497  computeResidual     residual = tile - upsample(parent approximation)
 *     :499-512  encodeResidual      short(roundf(residual)), little-endian bytes
 *     :514-559  computeApproxTile   approximation = upsample(parent approximation) + rounded residual
 *     :255-324  buildResiduals      the per-level loop calling the three in that order
 *
 * The upsample is the one ResidualProducer::upsample applies when it decodes (same taps, same CPU
 * evaluation order, ResidualProducer.cpp:342-384), which is what makes the approximation the builder
 * carries to the next level equal to what the run-time producer reconstructs.
 *
 * Row stride of all tiles is n = tileSize + 5 of the container (HeightMipmap.cpp:456), whatever the
 * tile size of the level.
 */
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <math.h>
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#include "orc.h"

/* the predicted height of texel (i, j) of tile (tx, ty): HeightMipmap.cpp:458-489 */
static float hm_predict(const float *parentTile, int n, int i, int j, int px, int py)
{
#define P(a, b) parentTile[(a) + (b) * n]
    float z;
    if (j % 2 == 0) {
        if (i % 2 == 0) {
            z = P(i / 2 + px, j / 2 + py);
        } else {
            float z0 = P(i / 2 + px - 1, j / 2 + py);
            float z1 = P(i / 2 + px, j / 2 + py);
            float z2 = P(i / 2 + px + 1, j / 2 + py);
            float z3 = P(i / 2 + px + 2, j / 2 + py);
            z = ((z1 + z2) * 9 - (z0 + z3)) / 16;
        }
    } else {
        if (i % 2 == 0) {
            float z0 = P(i / 2 + px, j / 2 - 1 + py);
            float z1 = P(i / 2 + px, j / 2 + py);
            float z2 = P(i / 2 + px, j / 2 + 1 + py);
            float z3 = P(i / 2 + px, j / 2 + 2 + py);
            z = ((z1 + z2) * 9 - (z0 + z3)) / 16;
        } else {
            z = 0;
            for (int dj = -1; dj <= 2; ++dj) {
                float f = dj == -1 || dj == 2 ? -1 / 16.0 : 9 / 16.0;
                for (int di = -1; di <= 2; ++di) {
                    float g = di == -1 || di == 2 ? -1 / 16.0 : 9 / 16.0;
                    z += f * g * P(i / 2 + di + px, j / 2 + dj + py);
                }
            }
        }
    }
#undef P
    return z;
}

/* One tile of one level.  parentTile: the approximation of the parent tile; tile: the heights of this
 * tile (already divided by the file's scale, HeightMipmap.cpp:410-418); both with row stride n.
 * Out: resid (dense (ts+5)^2 int16, the bytes encodeResidual writes), approx (row stride n),
 * maxR = max |residual before rounding|, maxErr = max |tile - approx|. */
void orc_hm_encode_tile(const float *parentTile, const float *tile, int n, int tileSize, int tx, int ty,
                        short *resid, float *approx, float *maxR, float *maxErr)
{
    const int px = 1 + (tx % 2) * tileSize / 2;
    const int py = 1 + (ty % 2) * tileSize / 2;
    float mr = 0.0f, me = 0.0f;
    for (int j = 0; j <= tileSize + 4; ++j) {
        for (int i = 0; i <= tileSize + 4; ++i) {
            const float z = hm_predict(parentTile, n, i, j, px, py);
            const int off = i + j * n;
            const float diff = tile[off] - z;                     /* computeResidual :492-494 */
            mr = fmaxf(diff < 0.0f ? -diff : diff, mr);
            const short q = (short) roundf(diff);                 /* encodeResidual :505 */
            const float r = (float) q;                            /* residual[off] = z (:506) */
            resid[i + j * (tileSize + 5)] = q;
            const float a = z + r;                                /* computeApproxTile :553-555 */
            const float err = tile[off] - a;
            me = fmaxf(err < 0.0f ? -err : err, me);
            approx[off] = a;
        }
    }
    if (maxR) *maxR = mr;
    if (maxErr) *maxErr = me;
}

/* ---------------------------------------------------------------------------------------------------
 * The height pyramid of a cube (the part of the builder BEFORE buildResiduals):
 *   HeightMipmap.cpp:67-81    setCube            which face lies left / right / below / above each face, and the
 *                                                rotation that maps coordinates across that edge
 *   HeightMipmap.cpp:149-254  buildBaseLevelTiles / buildMipmapLevel   level l = every second sample of level l + 1
 *                                                ((short) getTileHeight(2 i, 2 j): pure decimation, no averaging)
 *   HeightMipmap.cpp:327-372  getTileHeight      samples near a cube corner collapse onto the corner, samples past an
 *                                                edge come from the neighbouring face (rotated)
 *   AbstractTileCache.cpp:73-92                  clamp to [0, width] when there is no neighbour (flat DEMs)
 *   HeightMipmap.cpp:404-412  getTile            the (ts + 5)^2 height tile (tx, ty) of a level, / scale
 *   Preprocess.cpp:155-213, 406-445              the six cube projections and SphericalHeightFunction (lon / lat bilinear)
 * The mipmap tiles the reference keeps on disk are, sample for sample, the base level decimated: stored_l(x, y) =
 * base(x << (L - l), y << (L - l)) -- the stitched borders it also stores are never read back (getTileHeight reads
 * borders from the neighbour face instead) -- so the pyramid is restated as a gather from the six base grids.
 * ------------------------------------------------------------------------------------------------- */
static const int HM_NEIGH[6][4] = {      /* left, right, bottom, top of hm1..hm6 (0-based faces) */
    { 4, 2, 1, 3 }, { 4, 2, 5, 0 }, { 1, 3, 5, 0 }, { 2, 4, 5, 0 }, { 3, 1, 5, 0 }, { 4, 2, 3, 1 } };
static const int HM_ROT[6][4] = {        /* leftr, rightr, bottomr, topr */
    { 3, 1, 0, 2 }, { 0, 0, 0, 0 }, { 0, 0, 1, 3 }, { 0, 0, 2, 2 }, { 0, 0, 3, 1 }, { 1, 3, 2, 0 } };

static void hm_rotation(int r, int n, int x, int y, int *xp, int *yp)      /* ColorMipmap.cpp:421-441 */
{
    switch (r) {
    case 0: *xp = x; *yp = y; break;
    case 1: *xp = y; *yp = n - 1 - x; break;
    case 2: *xp = n - 1 - x; *yp = n - 1 - y; break;
    default: *xp = n - 1 - y; *yp = x; break;
    }
}

/* height sample (x, y) of `level` of face `face`; faces[f]: (B + 1)^2 base-level samples, row-major;
 * nfaces: 6 (a cube, setCube) or 1 (a flat DEM: no neighbours) */
float orc_hm_height(const short *const *faces, int nfaces, int B, int maxLevel, int level, int face, int x, int y)
{
    const int levelSize = 1 + (B >> (maxLevel - level));
    for (int hop = 0; hop < 8; ++hop) {
        const int cube = nfaces == 6;
        if (cube) {
            if (x <= 2 && y <= 2) { x = 0; y = 0; }
            else if (x > levelSize - 4 && y <= 2) { x = levelSize - 1; y = 0; }
            else if (x <= 2 && y > levelSize - 4) { x = 0; y = levelSize - 1; }
            else if (x > levelSize - 4 && y > levelSize - 4) { x = levelSize - 1; y = levelSize - 1; }
            int side = -1, ax = 0, ay = 0;
            if (x < 0) { side = 0; ax = levelSize - 1 + x; ay = y; }
            else if (x >= levelSize) { side = 1; ax = x - levelSize + 1; ay = y; }
            else if (y < 0) { side = 2; ax = x; ay = levelSize - 1 + y; }
            else if (y >= levelSize) { side = 3; ax = x; ay = y - levelSize + 1; }
            if (side >= 0) {
                hm_rotation(HM_ROT[face][side], levelSize, ax, ay, &x, &y);
                face = HM_NEIGH[face][side];
                continue;
            }
        }
        break;
    }
    /* AbstractTileCache::getTileHeight: clamp to [0, width], width = levelSize - 1 */
    const int w = levelSize - 1;
    x = x < 0 ? 0 : (x > w ? w : x);
    y = y < 0 ? 0 : (y > w ? w : y);
    const int sh = maxLevel - level;
    return (float) faces[face][((long) y << sh) * (B + 1) + ((long) x << sh)];
}

/* HeightMipmap::getTile: tile (tx, ty) of `level`, (ts + 5)^2 samples with row stride n = tileSize + 5 */
void orc_hm_get_tile(const short *const *faces, int nfaces, int B, int maxLevel, int topLevelSize, int tileSize, float scale,
                     int level, int face, int tx, int ty, float *tile)
{
    int ts = topLevelSize << level;
    if (ts > tileSize) ts = tileSize;
    for (int j = 0; j <= ts + 4; ++j)
        for (int i = 0; i <= ts + 4; ++i)
            tile[i + j * (tileSize + 5)] = orc_hm_height(faces, nfaces, B, maxLevel, level, face, i + ts * tx - 2, j + ts * ty - 2) / scale;
}

/* Preprocess.cpp:155-213: sample (x, y) of face 0..5 of a w-wide grid -> direction on the sphere */
void orc_cube_projection(int face, int x, int y, int w, double *sx, double *sy, double *sz)
{
    const double cx = x < 0 ? 0.0 : (x > w ? (double) w : (double) x), cy = y < 0 ? 0.0 : (y > w ? (double) w : (double) y);
    const double xl = cx / w * 2.0 - 1.0, yl = cy / w * 2.0 - 1.0;
    const double l = sqrt(xl * xl + yl * yl + 1.0);
    switch (face) {
    case 0: *sx = xl / l; *sy = yl / l; *sz = 1.0 / l; break;
    case 1: *sx = xl / l; *sy = -1.0 / l; *sz = yl / l; break;
    case 2: *sx = 1.0 / l; *sy = xl / l; *sz = yl / l; break;
    case 3: *sx = -xl / l; *sy = 1.0 / l; *sz = yl / l; break;
    case 4: *sx = -1.0 / l; *sy = -xl / l; *sz = yl / l; break;
    default: *sx = xl / l; *sy = -yl / l; *sz = -1.0 / l; break;
    }
}

/* SphericalHeightFunction::getHeight (Preprocess.cpp:420-444) + buildBaseLevelTile's (short) h: base-level sample
 * (x, y) of a face from an equirectangular source map src (sw x sh floats, row-major, row 0 = latitude 0 = north) */
short orc_spherical_base_sample(const float *src, int sw, int sh, int face, int x, int y, int B)
{
    double sx, sy, sz;
    orc_cube_projection(face, x, y, B, &sx, &sy, &sz);
    double lon = atan2(sy, sx) + M_PI;
    double lat = acos(sz);
    lon = lon / M_PI * (sw / 2);
    lat = lat / M_PI * sh;
    const int ilon = (int) floor(lon), ilat = (int) floor(lat);
    lon -= ilon;
    lat -= ilat;
    const double clon = 1.0 - lon, clat = 1.0 - lat;
    /* InputMap::get clamps nothing itself; the reference's maps are padded by one row: rows are clamped here */
    const int r0 = ilat < 0 ? 0 : (ilat > sh - 1 ? sh - 1 : ilat), r1 = ilat + 1 > sh - 1 ? sh - 1 : (ilat + 1 < 0 ? 0 : ilat + 1);
    const float h1 = src[(size_t) r0 * sw + (ilon + sw) % sw], h2 = src[(size_t) r0 * sw + (ilon + sw + 1) % sw];
    const float h3 = src[(size_t) r1 * sw + (ilon + sw) % sw], h4 = src[(size_t) r1 * sw + (ilon + sw + 1) % sw];
    const float h = (float) ((h1 * clon + h2 * lon) * clat + (h3 * clon + h4 * lon) * lat);
    return (short) h;
}

/* all (B + 1)^2 base-level samples of one face, row-major (buildBaseLevelTiles over the interior of every tile) */
void orc_spherical_base_grid(const float *src, int sw, int sh, int face, int B, short *out)
{
    for (int y = 0; y <= B; ++y)
        for (int x = 0; x <= B; ++x)
            out[(size_t) y * (B + 1) + x] = orc_spherical_base_sample(src, sw, sh, face, x, y, B);
}

/* PlaneHeightFunction::getHeight (Preprocess.cpp:335-366) + (short) h: the base grid of a flat DEM (preprocessDem) */
void orc_plane_base_grid(const float *src, int sw, int sh, int B, short *out)
{
    for (int gy = 0; gy <= B; ++gy)
        for (int gx = 0; gx <= B; ++gx) {
            double x = (double) gx / B * sw, y = (double) gy / B * sh;
            const int ix = (int) floor(x), iy = (int) floor(y);
            x -= ix;
            y -= iy;
            const double cx = 1.0 - x, cy = 1.0 - y;
#define ORC_CL(v, n) ((v) < 0 ? 0 : ((v) > (n) - 1 ? (n) - 1 : (v)))
            const float h1 = src[(size_t) ORC_CL(iy, sh) * sw + ORC_CL(ix, sw)], h2 = src[(size_t) ORC_CL(iy, sh) * sw + ORC_CL(ix + 1, sw)];
            const float h3 = src[(size_t) ORC_CL(iy + 1, sh) * sw + ORC_CL(ix, sw)], h4 = src[(size_t) ORC_CL(iy + 1, sh) * sw + ORC_CL(ix + 1, sw)];
#undef ORC_CL
            const float h = (float) ((h1 * cx + h2 * x) * cy + (h3 * cx + h4 * x) * y);
            out[(size_t) gy * (B + 1) + gx] = (short) h;
        }
}
