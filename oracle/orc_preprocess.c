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
#include <math.h>

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
