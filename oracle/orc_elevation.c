/*
 * orc_elevation.c -- ORACLE (test infrastructure, never shipped): the
 * ElevationProducer pass.
 *
 * Restates:
 *   terrain/sources/proland/dem/ElevationProducer.cpp:293-376 (uniform block)
 *   src/demo/shaders/elevation/upsampleShader.glsl:28-203    (variant D)
 *   src/terrain/examples/terrain{1,2,4}/upsampleShader.glsl  (variants A,B,C)
 *   terrain/sources/proland/dem/CPUElevationProducer.cpp:194-245
 *   core/sources/proland/terrain/TileSamplerZ.cpp:43-133 (min/max region)
 *
 * Evaluation follows the canonical fp32 order of orc_fp.h (dot products are
 * left-to-right fma chains, a*b+c is one fma; -DORC_STRICT: no contraction),
 * mat4 sums in the order GLSL's `mdot` spells them.
 * Texture fetches are NEAREST on exact texel centres (the shader's own
 * coordinate maths resolves to the integer texels used here; checked by
 * tests/test_oracle.py::test_texcoord_identities) with CLAMP_TO_EDGE.
 */
#include "orc.h"
#include "orc_fp.h"
#include <math.h>
#include <string.h>

/* --- GLSL constant matrices, upsampleShader.glsl:56-134, column-major:
 *     M[i][c][r] is column c, row r of matrix i (i = parity index). -------- */
typedef float mat4[4][4];

static const mat4 SLOPEX[4] = {
    { {0, 0, 0, 0}, {1.0f, 0, -1.0f, 0}, {0, 0, 0, 0}, {0, 0, 0, 0} },
    { {0, 0, 0, 0}, {0.5f, 0.5f, -0.5f, -0.5f}, {0, 0, 0, 0}, {0, 0, 0, 0} },
    { {0, 0, 0, 0}, {0.5f, 0, -0.5f, 0}, {0.5f, 0, -0.5f, 0}, {0, 0, 0, 0} },
    { {0, 0, 0, 0}, {0.25f, 0.25f, -0.25f, -0.25f}, {0.25f, 0.25f, -0.25f, -0.25f}, {0, 0, 0, 0} },
};
static const mat4 SLOPEY[4] = {
    { {0, 1.0f, 0, 0}, {0, 0, 0, 0}, {0, -1.0f, 0, 0}, {0, 0, 0, 0} },
    { {0, 0.5f, 0.5f, 0}, {0, 0, 0, 0}, {0, -0.5f, -0.5f, 0}, {0, 0, 0, 0} },
    { {0, 0.5f, 0, 0}, {0, 0.5f, 0, 0}, {0, -0.5f, 0, 0}, {0, -0.5f, 0, 0} },
    { {0, 0.25f, 0.25f, 0}, {0, 0.25f, 0.25f, 0}, {0, -0.25f, -0.25f, 0}, {0, -0.25f, -0.25f, 0} },
};
static const mat4 CURV[4] = {
    { {0, -1.0f, 0, 0}, {-1.0f, 4.0f, -1.0f, 0}, {0, -1.0f, 0, 0}, {0, 0, 0, 0} },
    { {0, -0.5f, -0.5f, 0}, {-0.5f, 1.5f, 1.5f, -0.5f}, {0, -0.5f, -0.5f, 0}, {0, 0, 0, 0} },
    { {0, -0.5f, 0, 0}, {-0.5f, 1.5f, -0.5f, 0}, {-0.5f, 1.5f, -0.5f, 0}, {0, -0.5f, 0, 0} },
    { {0, -0.25f, -0.25f, 0}, {-0.25f, 0.5f, 0.5f, -0.25f}, {-0.25f, 0.5f, 0.5f, -0.25f}, {0, -0.25f, -0.25f, 0} },
};
#define W1 (-1.0f / 16.0f)
#define W9 (9.0f / 16.0f)
#define V1 (1.0f / 256.0f)
#define V9 (-9.0f / 256.0f)
#define V81 (81.0f / 256.0f)
static const mat4 UPSAMPLE[4] = {
    { {0, 0, 0, 0}, {0, 1.0f, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0} },
    { {0, 0, 0, 0}, {W1, W9, W9, W1}, {0, 0, 0, 0}, {0, 0, 0, 0} },
    { {0, W1, 0, 0}, {0, W9, 0, 0}, {0, W9, 0, 0}, {0, W1, 0, 0} },
    { {V1, V9, V9, V1}, {V9, V81, V81, V9}, {V9, V81, V81, V9}, {V1, V9, V9, V1} },
};

/* GLSL dot(vec4,vec4): canonical fma chain (orc_fp.h R1) */
static inline float dot4(const float *a, const float *b) { return orc_dot4(a, b); }
/* upsampleShader.glsl:136-138 */
static inline float mdot(const mat4 a, const mat4 b)
{
    return dot4(a[0], b[0]) + dot4(a[1], b[1]) + dot4(a[2], b[2]) + dot4(a[3], b[3]);
}
static inline float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
static inline int clampi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }
static inline int floordiv(int a, int b) { int q = a / b; return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q; }
static inline int posmod(int a, int b) { int m = a % b; return m < 0 ? m + b : m; }

/* parent texel fetch, channel ch, CLAMP_TO_EDGE */
static inline float ptexel(const float *parent, int W, int x, int y, int ch)
{
    x = clampi(x, 0, W - 1);
    y = clampi(y, 0, W - 1);
    return parent[(x + y * W) * 3 + ch];
}

/* ElevationProducer.cpp:293-376 */
void orc_elev_uniforms(int W, int gridMeshSize, float rootQuadSize, int flip,
                       const float *noiseAmp, int nAmp, int face,
                       int level, int tx, int ty,
                       int has_resid, int resid_W,
                       int noise_mode, int no_clamp, orc_elev_params *p)
{
    const int tileSize = W - 5;
    memset(p, 0, sizeof(*p));
    p->W = W;
    p->level = level;
    /* getRootQuadSize() / (1 << level) / tileSize : float / int / int */
    p->pixel_size = rootQuadSize / (float) (1 << level) / (float) tileSize;
    p->grid = (W - 5) / gridMeshSize;
    p->flip = flip;
    p->dx = (tx % 2) * (tileSize / 2);
    p->dy = (ty % 2) * (tileSize / 2);
    p->has_resid = has_resid;
    if (has_resid) {
        int mod = (resid_W - 5) / tileSize;
        p->rx = (tx % mod) * tileSize;
        p->ry = (ty % mod) * tileSize;
        p->resid_stride = resid_W;
    }
    p->rs = level < nAmp ? noiseAmp[level] : 0.0f;
    orc_noise_select(level, tx, ty, face, &p->noiseR, &p->noiseL);
    p->noise_mode = noise_mode;
    p->no_clamp = no_clamp;
}

/* upsampleShader.glsl `main` */
void orc_upsample_tile(const orc_elev_params *p, const float *parent,
                       const float *resid, const float *noise, float *out)
{
    const int W = p->W;
    const int g = p->grid;
    const float *nlayer = noise + (size_t) p->noiseL * W * W;

    for (int y = 0; y < W; ++y) {
        for (int x = 0; x < W; ++x) {
            /* residual: texel (x + 0.25) of the uploaded W x W window, which is
             * the producer's copy of resid[(x+rx),(y+ry)] */
            float zf = 0.0f;
            if (p->has_resid && resid != NULL) {
                zf = 1.0f * resid[(x + p->rx) + (y + p->ry) * p->resid_stride];
            }

            /* 4x4 parent neighbourhood, cz[c][r] = parent.zf[bx + r, by + c] */
            mat4 cz;
            const int bx = (x >> 1) + p->dx;
            const int by = (y >> 1) + p->dy;
            for (int c = 0; c < 4; ++c) {
                for (int r = 0; r < 4; ++r) {
                    cz[c][r] = parent ? ptexel(parent, W, bx + r, by + c, 0) : 0.0f;
                }
            }
            const int i = (x & 1) + 2 * (y & 1);

            /* noise texel under rotation noiseR: uvs = (nx, ny, 1-nx, 1-ny),
             * components noiseR and (noiseR+1)%4 */
            int nxi, nyi;
            switch (p->noiseR) {
            case 0: nxi = x; nyi = y; break;
            case 1: nxi = y; nyi = W - 1 - x; break;
            case 2: nxi = W - 1 - x; nyi = W - 1 - y; break;
            default: nxi = W - 1 - y; nyi = x; break;
            }
            const float n = nlayer[nxi + nyi * W];

            if (p->noise_mode == ORC_NOISE_PLAIN) {
                zf = orc_fma(fabsf(p->rs), n, zf);
            } else {
                float nvx = mdot(cz, SLOPEX[i]);
                float nvy = mdot(cz, SLOPEY[i]);
                float nvz = 2.0f * p->pixel_size;
                float nv[2] = { nvx, nvy };
                float slope = sqrtf(orc_dot2(nv, nv)) / nvz;   /* length(n.xy) / n.z */
                float curvature = mdot(cz, CURV[i]) / p->pixel_size;
                float amp = fmaxf(clampf(4.0f * curvature, 0.0f, 1.5f),
                                  clampf(orc_fma(2.0f, slope, -0.5f), 0.1f, 4.0f));
                if (p->rs < 0.0f) {
                    zf = orc_fma(-p->rs, n, zf);
                } else {
                    zf = orc_fma(amp * p->rs, n, zf);
                }
            }

            float zc = zf;
            if (parent != NULL) {
                zf = zf + mdot(cz, UPSAMPLE[i]);

                /* coarse (parent-mesh) height at this vertex */
                const int ix = x - 2, iy = y - 2;
                const int kx_round = floordiv(ix + g, 2 * g);   /* floor(ij.x/2g + 0.5) */
                const int kx_floor = floordiv(ix, 2 * g);
                const int ky_round = floordiv(iy + g, 2 * g);
                const int ky_floor = floordiv(iy, 2 * g);
                const int ux = 2 + g * kx_round + p->dx;   /* uvc.x */
                const int uy = 2 + g * ky_floor + p->dy;   /* uvc.y */
                const int vx = 2 + g * kx_floor + p->dx;   /* uvc.z */
                const int vy = 2 + g * ky_round + p->dy;   /* uvc.w */
                float zc1 = ptexel(parent, W, ux, uy, 2);
                float zc3 = ptexel(parent, W, vx, vy, 2);
                if (p->flip && posmod(ix, 2 * g) == g && posmod(iy, 2 * g) == g) {
                    float zc0 = ptexel(parent, W, vx, uy, 2);
                    float zc2 = ptexel(parent, W, ux, vy, 2);
                    zc = (zc3 + zc1 >= zc0 + zc2 ? zc1 + zc3 : zc0 + zc2) * 0.5f;
                } else {
                    zc = (zc1 + zc3) * 0.5f;
                }
            }

            float *o = out + (size_t) (x + y * W) * 3;
            o[0] = zf;
            o[1] = zc;
            o[2] = p->no_clamp ? zf : fmaxf(zf, 0.0f);
        }
    }
}

/* CPUElevationProducer.cpp:194-245: the reference's own CPU statement of the
 * upsample + residual arithmetic (single channel, no noise).  Kept in the
 * reference's evaluation order: ((z1+z2)*9-(z0+z3))/16 and the 16-term
 * accumulation z += f*g*parent. */
void orc_cpu_elevation_tile(int W, int level, int tx, int ty, const float *parent,
                            const float *resid, int resid_W, int rx, int ry, float *out)
{
    const int tileSize = W - 5;
    const int px = 1 + (tx % 2) * tileSize / 2;
    const int py = 1 + (ty % 2) * tileSize / 2;
    for (int j = 0; j < W; ++j) {
        for (int i = 0; i < W; ++i) {
            float z;
            if (level == 0) {
                z = 0.0f;
            } else {
                const int cx = i / 2 + px, cy = j / 2 + py;
#define P(a, b) parent[(a) + (b) * W]
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
                        z = 0;
                        for (int dj = -1; dj <= 2; ++dj) {
                            float f = (dj == -1 || dj == 2) ? -1 / 16.0f : 9 / 16.0f;
                            for (int di = -1; di <= 2; ++di) {
                                float g = (di == -1 || di == 2) ? -1 / 16.0f : 9 / 16.0f;
                                z += f * g * P(cx + di, cy + dj);
                            }
                        }
                    }
                }
#undef P
            }
            float r = resid ? resid[(i + rx) + (j + ry) * resid_W] : 0.0f;
            out[i + j * W] = z + r;
        }
    }
}

/* TileSamplerZ.cpp:60-64: texels 2.5 .. W-2.5, i.e. [2, W-3]^2 of channel z */
void orc_tile_minmax(int W, const float *elev, float *zmin, float *zmax)
{
    float lo = INFINITY, hi = -INFINITY;
    for (int y = 2; y <= W - 3; ++y) {
        for (int x = 2; x <= W - 3; ++x) {
            float z = elev[(x + y * W) * 3 + 2];
            lo = fminf(lo, z);
            hi = fmaxf(hi, z);
        }
    }
    *zmin = lo;
    *zmax = hi;
}
