/*
 * pl_reqmath.cuh -- the per-tile "uniform" maths of the two producers, written
 * once and compiled for BOTH the host (pl_elev_make_req / pl_norm_make_req, the
 * per-tile plugin path) and the device (pl_requests.cu: whole Morton ranges of
 * tiles get their requests generated on the GPU, so a full-quadtree sweep never
 * waits for the CPU).
 *
 *   2-D cnoise                    core/sources/proland/math/noise.cpp:117-165
 *   noise layer / rotation select terrain/sources/proland/dem/ElevationProducer.cpp:345-373
 *   per-tile elevation uniforms   terrain/sources/proland/dem/ElevationProducer.cpp:305-343
 *   per-tile normal uniforms      terrain/sources/proland/dem/NormalProducer.cpp:196-283
 *
 * No contraction on either side (--fmad=false / -ffp-contract=off): cnoise
 * decides integers and the fp64 patch geometry is narrowed to fp32, so host and
 * device must round identically.  The only fused operation is the explicit
 * fmaf() in the smoothstep polynomial (canonical order, oracle/orc_fp.h).
 */
#ifndef PL_REQMATH_CUH
#define PL_REQMATH_CUH

#include <math.h>
#include "pl_internal.h"

#ifdef __CUDACC__
#define PL_HD __host__ __device__ __forceinline__
#else
#define PL_HD inline
#endif

constexpr int kPerlinB = 256;
constexpr int kPerlinN = 2 * kPerlinB + 2;

struct PerlinView {
    const int *perm;     /* kPerlinN */
    const float *g2;     /* kPerlinN x 2 */
};

struct PerlinAxis { int b0, b1; float r0, r1; };

PL_HD PerlinAxis perlin_split(float v)
{
    const float t = v + 4096.0f;
    PerlinAxis a;
    a.b0 = ((int) t) & (kPerlinB - 1);            /* truncation, like the reference */
    a.b1 = (a.b0 + 1) & (kPerlinB - 1);
    a.r0 = t - (float) (int) floorf(t);           /* ... but floor for the fraction */
    a.r1 = a.r0 - 1.0f;
    return a;
}
PL_HD float perlin_fade(float t) { return t * t * (3.0f - 2.0f * t); }
PL_HD float perlin_mix(float t, float a, float b) { return a + t * (b - a); }

PL_HD float cnoise2(const PerlinView T, float x, float y)
{
    const PerlinAxis ax = perlin_split(x), ay = perlin_split(y);
    const int i = T.perm[ax.b0], j = T.perm[ax.b1];
    const int b00 = T.perm[i + ay.b0], b10 = T.perm[j + ay.b0];
    const int b01 = T.perm[i + ay.b1], b11 = T.perm[j + ay.b1];
    const float sx = perlin_fade(ax.r0), sy = perlin_fade(ay.r0);
    float u = ax.r0 * T.g2[2 * b00] + ay.r0 * T.g2[2 * b00 + 1];
    float v = ax.r1 * T.g2[2 * b10] + ay.r0 * T.g2[2 * b10 + 1];
    const float lo = perlin_mix(sx, u, v);
    u = ax.r0 * T.g2[2 * b01] + ay.r1 * T.g2[2 * b01 + 1];
    v = ax.r1 * T.g2[2 * b11] + ay.r1 * T.g2[2 * b11 + 1];
    const float hi = perlin_mix(sx, u, v);
    return perlin_mix(sy, lo, hi);
}

/* sign test of cnoise at a lattice point given as (double, double): the C++
 * call narrows the double arguments to the float parameters */
PL_HD int noise_bit(const PerlinView T, double a, double b) { return cnoise2(T, (float) a, (float) b) > 0.0f ? 1 : 0; }

/* bit0 bottom, bit1 right, bit2 top, bit3 left; the six cube faces are unfolded
 * into one integer lattice so the shared edge of two tiles gets the same bit on
 * both sides; face 0 (flat terrain) takes the generic branch */
PL_HD void noise_select(const PerlinView T, int level, int tx, int ty, int face, int *noiseR, int *noiseL)
{
    const int n = 1 << level;
    int bottom, right, top, left;
    if (face == 1) {
        bottom = noise_bit(T, tx + 0.5, ty + n);
        right = tx == n - 1 ? noise_bit(T, ty + n + 0.5, n) : noise_bit(T, tx + 1, ty + n + 0.5);
        top = ty == n - 1 ? noise_bit(T, (3 * n - 1 - tx) + 0.5, n) : noise_bit(T, tx + 0.5, ty + n + 1);
        left = tx == 0 ? noise_bit(T, (4 * n - 1 - ty) + 0.5, n) : noise_bit(T, tx, ty + n + 0.5);
    } else if (face == 6) {
        bottom = ty == 0 ? noise_bit(T, (3 * n - 1 - tx) + 0.5, 0) : noise_bit(T, tx + 0.5, ty - n);
        right = tx == n - 1 ? noise_bit(T, (2 * n - 1 - ty) + 0.5, 0) : noise_bit(T, tx + 1, ty - n + 0.5);
        top = noise_bit(T, tx + 0.5, ty - n + 1);
        left = tx == 0 ? noise_bit(T, 3 * n + ty + 0.5, 0) : noise_bit(T, tx, ty - n + 0.5);
    } else {
        const int off = n * (face - 2);
        bottom = noise_bit(T, tx + off + 0.5, ty);
        right = noise_bit(T, (tx + off + 1) % (4 << level), ty + 0.5);   /* C remainder, may be < 0 */
        top = noise_bit(T, tx + off + 0.5, ty + 1);
        left = noise_bit(T, tx + off, ty + 0.5);
    }
    const int bits = bottom | (right << 1) | (top << 2) | (left << 3);
    /* rotation / layer that puts the two border seeds where the bits say */
    *noiseR = (int) ((0x1ADF1210u >> (2 * bits)) & 3u);           /* 0,0,1,0,2,0,1,0,3,3,1,3,2,2,1,0 */
    *noiseL = (int) ((0x5442432142312110ull >> (4 * bits)) & 15u); /* 0,1,1,2,1,3,2,4,1,2,3,4,2,4,4,5 */
}

PL_HD void elev_fill_req(const PerlinView T, int tile_w, float root_quad_size, const float *noise_amp, int n_amp,
                         int face, int level, int tx, int ty, int resid_tile_w, int has_resid, pl_elev_req *req)
{
    const int tileSize = tile_w - 5;
    req->out_slot = -1;
    req->parent_slot = -1;
    req->resid_slot = -1;
    req->dx = (tx % 2) * (tileSize / 2);
    req->dy = (ty % 2) * (tileSize / 2);
    req->rx = 0;
    req->ry = 0;
    if (has_resid && resid_tile_w > 0) {
        const int mod = (resid_tile_w - 5) / tileSize;
        req->rx = (tx % mod) * tileSize;
        req->ry = (ty % mod) * tileSize;
    }
    req->rs = level < n_amp ? noise_amp[level] : 0.0f;
    /* float / int / int, as getRootQuadSize() / (1 << level) / tileSize evaluates */
    req->pixel_size = root_quad_size / (float) (1 << level) / (float) tileSize;
    noise_select(T, level, tx, ty, face, &req->noise_r, &req->noise_l);
    req->level = level;
    req->tx = tx;
    req->ty = ty;
    req->pad_[0] = req->pad_[1] = 0;
}

/* the uniforms of one ortho tile, OrthoProducer.cpp:286-366 (slots are left at -1) */
PL_HD void ortho_fill_req(const PerlinView T, const pl_ortho_scene *sc, int level, int tx, int ty, pl_ortho_req *q)
{
    const int half = (sc->tile_w - 4) / 2;
    q->out_slot = -1;
    q->parent_slot = -1;
    q->resid_slot = -1;
    q->dx = (tx % 2) * half;
    q->dy = (ty % 2) * half;
    noise_select(T, level, tx, ty, sc->face, &q->noise_r, &q->noise_l);
    q->level = level;
    const float rs = level < sc->n_amp ? sc->noise_amp[level] : 0.0f;
    if (sc->hsv) {   /* noiseColor * (rs, rs, rs, scale * rs) / 255 */
        for (int c = 0; c < 3; ++c) q->noise_color[c] = sc->noise_color[c] * rs / 255.0f;
        q->noise_color[3] = sc->noise_color[3] * (sc->scale * rs) / 255.0f;
    } else {         /* noiseColor * scale * rs / 255 */
        for (int c = 0; c < 4; ++c) q->noise_color[c] = sc->noise_color[c] * sc->scale * rs / 255.0f;
    }
    q->tx = tx;
    q->ty = ty;
    q->pad_[0] = q->pad_[1] = 0;
}

struct PlV3 { double x, y, z; };
PL_HD PlV3 v3_unit(PlV3 v, double *len)
{
    const double l = sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
    const double inv = 1.0 / l;
    if (len) *len = l;
    PlV3 r = { v.x * inv, v.y * inv, v.z * inv };
    return r;
}
PL_HD PlV3 v3_cross(PlV3 a, PlV3 b)
{
    PlV3 r = { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x };
    return r;
}
/* rows of the world -> tangent frame at the cube-face point (px, py, R) */
PL_HD void v3_frame(double px, double py, double R, PlV3 *ux, PlV3 *uy, PlV3 *uz)
{
    const PlV3 pc = { px, py, R }, unit_y = { 0.0, 1.0, 0.0 };
    *uz = v3_unit(pc, nullptr);
    *ux = v3_unit(v3_cross(unit_y, *uz), nullptr);
    *uy = v3_cross(*uz, *ux);
}

PL_HD void norm_fill_req(int sphere, double root_quad_size, int level, int tx, int ty, pl_norm_req *req)
{
    req->out_slot = req->elev_slot = -1;
    req->parent_slot = -1;
    req->ptx = tx % 2;
    req->pty = ty % 2;
    req->level = level;
    req->pad_[0] = req->pad_[1] = req->pad_[2] = 0;

    const double D = root_quad_size, R = D / 2.0;
    const double n = (double) (1 << level);
    const double x0 = (double) tx / n * D - R, y0 = (double) ty / n * D - R;
    req->deform[0] = (float) x0;
    req->deform[1] = (float) y0;
    req->deform[2] = (float) (D / n);
    for (int k = 0; k < 9; ++k) req->w2t[k] = req->p2t[k] = (k % 4 == 0) ? 1.0f : 0.0f;
    for (int k = 0; k < 12; ++k) req->corners[k] = req->verticals[k] = 0.0f;
    for (int k = 0; k < 4; ++k) req->norms[k] = 0.0f;
    req->smooth = 0.0f;
    if (!sphere) {
        req->deform[3] = 0.0f;
        return;
    }
    req->deform[3] = (float) R;

    const double x1 = (double) (tx + 1) / n * D - R, y1 = (double) (ty + 1) / n * D - R;
    const PlV3 corner[4] = { { x0, y0, R }, { x1, y0, R }, { x0, y1, R }, { x1, y1, R } };
    PlV3 v[4];
    double len[4];
    for (int k = 0; k < 4; ++k) v[k] = v3_unit(corner[k], &len[k]);
    const PlV3 vc = { (v[0].x + v[1].x + v[2].x + v[3].x) * 0.25, (v[0].y + v[1].y + v[2].y + v[3].y) * 0.25,
                      (v[0].z + v[1].z + v[2].z + v[3].z) * 0.25 };
    for (int k = 0; k < 4; ++k) {
        req->corners[0 + k] = (float) (v[k].x * R - vc.x * R);
        req->corners[4 + k] = (float) (v[k].y * R - vc.y * R);
        req->corners[8 + k] = (float) (v[k].z * R - vc.z * R);
        req->verticals[0 + k] = (float) v[k].x;
        req->verticals[4 + k] = (float) v[k].y;
        req->verticals[8 + k] = (float) v[k].z;
        req->norms[k] = (float) len[k];
    }
    PlV3 ux, uy, uz;
    v3_frame((x0 + x1) * 0.5, (y0 + y1) * 0.5, R, &ux, &uy, &uz);
    const double w2t[9] = { ux.x, ux.y, ux.z, uy.x, uy.y, uy.z, uz.x, uz.y, uz.z };
    for (int k = 0; k < 9; ++k) req->w2t[k] = (float) w2t[k];
    if (level > 0) {
        const double np = (double) (1 << (level - 1));
        PlV3 pux, puy, puz;
        v3_frame((tx / 2 + 0.5) / np * D - R, (ty / 2 + 0.5) / np * D - R, R, &pux, &puy, &puz);
        const double t2w[9] = { pux.x, puy.x, puz.x, pux.y, puy.y, puz.y, pux.z, puy.z, puz.z };
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                req->p2t[i * 3 + j] = (float) (w2t[i * 3 + 0] * t2w[0 * 3 + j] + w2t[i * 3 + 1] * t2w[1 * 3 + j]
                                               + w2t[i * 3 + 2] * t2w[2 * 3 + j]);
    }
    /* smoothstep(R/32, R/64, deform.z) is tile-uniform: evaluated once per tile,
     * in fp32 and in the canonical order the oracle uses per texel */
    const float Rf = req->deform[3];
    const float e0 = Rf / 32.0f, e1 = Rf / 64.0f;
    float t = (req->deform[2] - e0) / (e1 - e0);
    t = fminf(fmaxf(t, 0.0f), 1.0f);
    req->smooth = t * t * fmaf(-2.0f, t, 3.0f);
}

/* does EVERY tile of `level` qualify for the register form of the FAST normal pass (plnorm::normal_reg_ok)?  Spheres: the
 * smoothstep factor of norm_fill_req is exactly 1 (it depends on the level only); flat scenes: the tangent frame norm_fill_req
 * writes is the identity.  The host decides per launch with this; the expressions are norm_fill_req's. */
PL_HD bool norm_level_all_reg(int sphere, double root_quad_size, int level)
{
    if (!sphere) return true;
    const double D = root_quad_size, R = D / 2.0;
    const float Rf = (float) R, dz = (float) (D / (double) (1 << level));
    const float e0 = Rf / 32.0f, e1 = Rf / 64.0f;
    float t = (dz - e0) / (e1 - e0);
    t = fminf(fmaxf(t, 0.0f), 1.0f);
    return t * t * fmaf(-2.0f, t, 3.0f) == 1.0f;
}

/* Morton (Z-order) index <-> (tx, ty): the four children of a quad are
 * (2tx,2ty), (2tx+1,2ty), (2tx,2ty+1), (2tx+1,2ty+1) in that order
 * (TileSampler.cpp:416-461), i.e. x in the even bits */
PL_HD uint32_t morton_compact(uint64_t v)
{
    v &= 0x5555555555555555ull;
    v = (v | (v >> 1)) & 0x3333333333333333ull;
    v = (v | (v >> 2)) & 0x0F0F0F0F0F0F0F0Full;
    v = (v | (v >> 4)) & 0x00FF00FF00FF00FFull;
    v = (v | (v >> 8)) & 0x0000FFFF0000FFFFull;
    v = (v | (v >> 16)) & 0x00000000FFFFFFFFull;
    return (uint32_t) v;
}
PL_HD void morton_decode(uint64_t m, int *tx, int *ty)
{
    *tx = (int) morton_compact(m);
    *ty = (int) morton_compact(m >> 1);
}

#endif
