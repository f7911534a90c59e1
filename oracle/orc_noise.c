/*
 * orc_noise.c -- ORACLE (test infrastructure, never shipped): the pseudo-random
 * and Perlin pieces of the elevation path.
 *
 * Restates:
 *   core/sources/proland/math/noise.h:50-67        (LCG, frandom)
 *   core/sources/proland/math/noise.cpp:68-165     (gradient tables, 2D cnoise)
 *   terrain/sources/proland/dem/ElevationProducer.cpp:50-128 (createDemNoise)
 *   terrain/sources/proland/dem/ElevationProducer.cpp:345-373 (layer select)
 *
 * Compile with -ffp-contract=off: cnoise feeds an integer decision and must
 * evaluate in plain IEEE fp32 (what an x86-64 SSE build of the reference does).
 */
#include "orc.h"
#include <math.h>
#include <string.h>

/* noise.h:50-54. 31-bit LCG; `long` is 64-bit here, the mask makes the result
 * identical to the reference's 32-bit-long Win32 build. */
long orc_lrandom(long *seed)
{
    long s = (*seed * 1103515245L + 12345L) & 0x7FFFFFFFL;
    *seed = s;
    return s;
}

/* noise.h:63-67: top 24 of the 31 bits, scaled to [0,1) */
float orc_frandom(long *seed)
{
    long top = orc_lrandom(seed) >> 7;
    return (float) top / (float) (1 << 24);
}

/* ------------------------------------------------------------------------
 * Classic Perlin tables, noise.cpp:68-101.  Only the 2D gradients are used by
 * the elevation path but the LCG is shared by g1/g2/g3, so all draws are made.
 * ------------------------------------------------------------------------ */
#define NB 256
static int perm[2 * NB + 2];
static float grad2[2 * NB + 2][2];
static int tables_ready = 0;

static float draw_unit(long *seed)
{
    /* (lrandom % 512 - 256) / 256 */
    return (float) ((orc_lrandom(seed) % (2 * NB)) - NB) / NB;
}

static void build_tables(void)
{
    long seed = 12345;
    int i;
    for (i = 0; i < NB; ++i) {
        perm[i] = i;
        (void) draw_unit(&seed);                /* g1[i] */
        float gx = draw_unit(&seed);
        float gy = draw_unit(&seed);
        float len = sqrtf(gx * gx + gy * gy);
        grad2[i][0] = gx / len;
        grad2[i][1] = gy / len;
        (void) draw_unit(&seed);                /* g3[i][0..2] */
        (void) draw_unit(&seed);
        (void) draw_unit(&seed);
    }
    /* shuffle: i = 255 .. 1 */
    for (i = NB - 1; i > 0; --i) {
        int j = (int) (orc_lrandom(&seed) % NB);
        int k = perm[i];
        perm[i] = perm[j];
        perm[j] = k;
    }
    for (i = 0; i < NB + 2; ++i) {
        perm[NB + i] = perm[i];
        grad2[NB + i][0] = grad2[i][0];
        grad2[NB + i][1] = grad2[i][1];
    }
    tables_ready = 1;
}

void orc_cnoise_tables(int *p, float *g2)
{
    if (!tables_ready) build_tables();
    memcpy(p, perm, sizeof(perm));
    memcpy(g2, grad2, sizeof(grad2));
}

static inline float s_curve(float t) { return t * t * (3.0f - 2.0f * t); }
static inline float lerpf(float t, float a, float b) { return a + t * (b - a); }

/* noise.cpp:117-165 with period == 0 */
float orc_cnoise2(float x, float y)
{
    if (!tables_ready) build_tables();

    float t = x + 4096.0f;
    int bx0 = ((int) t) & 0xFF;
    int bx1 = (bx0 + 1) & 0xFF;
    float rx0 = t - (float) (int) floorf(t);
    float rx1 = rx0 - 1.0f;

    t = y + 4096.0f;
    int by0 = ((int) t) & 0xFF;
    int by1 = (by0 + 1) & 0xFF;
    float ry0 = t - (float) (int) floorf(t);
    float ry1 = ry0 - 1.0f;

    int i = perm[bx0];
    int j = perm[bx1];
    int b00 = perm[i + by0];
    int b10 = perm[j + by0];
    int b01 = perm[i + by1];
    int b11 = perm[j + by1];

    float sx = s_curve(rx0);
    float sy = s_curve(ry0);

    float u = rx0 * grad2[b00][0] + ry0 * grad2[b00][1];
    float v = rx1 * grad2[b10][0] + ry0 * grad2[b10][1];
    float a = lerpf(sx, u, v);

    u = rx0 * grad2[b01][0] + ry1 * grad2[b01][1];
    v = rx1 * grad2[b11][0] + ry1 * grad2[b11][1];
    float b = lerpf(sx, u, v);

    return lerpf(sy, a, b);
}

/* ------------------------------------------------------------------------
 * createDemNoise, ElevationProducer.cpp:50-128.
 * Six layers; layer nl carries border pattern bits layers[nl] (bit0 bottom,
 * bit1 right, bit2 top, bit3 left).  Each border is drawn from its own LCG
 * stream (seed A or B by the bit) and mirrored so adjacent tiles agree; the
 * centre comes from ONE stream that keeps running across layers.
 * ------------------------------------------------------------------------ */
#define SEED_A 7654321L
#define SEED_B 5647381L

static inline float signed_draw(long *s) { return orc_frandom(s) * 2.0f - 1.0f; }

void orc_dem_noise(int W, float *out)
{
    static const int pattern[6] = { 0, 1, 3, 5, 7, 15 };
    long centre_seed = 1234567;
    const int half = W / 2;

    for (int nl = 0; nl < 6; ++nl) {
        float *n = out + (size_t) nl * W * W;
        const int bits = pattern[nl];
        long bs;
        int h, v;
#define PUT(x, y, val) n[(x) + (y) * W] = (val)

        memset(n, 0, sizeof(float) * W * W);

        /* bottom: row 2 symmetric about the middle, then rows 3,4 mirrored
         * into rows 1,0 with x reversed */
        bs = (bits & 1) ? SEED_B : SEED_A;
        for (h = 5; h <= half; ++h) {
            float r = signed_draw(&bs);
            PUT(h, 2, r);
            PUT(W - 1 - h, 2, r);
        }
        for (v = 3; v < 5; ++v) {
            for (h = 5; h < W - 5; ++h) {
                float r = signed_draw(&bs);
                PUT(h, v, r);
                PUT(W - 1 - h, 4 - v, r);
            }
        }

        /* right: column W-3, then columns W-4, W-5 mirrored into W-2, W-1 */
        bs = (bits & 2) ? SEED_B : SEED_A;
        for (v = 5; v <= half; ++v) {
            float r = signed_draw(&bs);
            PUT(W - 3, v, r);
            PUT(W - 3, W - 1 - v, r);
        }
        for (h = W - 4; h >= W - 5; --h) {
            for (v = 5; v < W - 5; ++v) {
                float r = signed_draw(&bs);
                PUT(h, v, r);
                PUT(2 * W - 6 - h, W - 1 - v, r);
            }
        }

        /* top: row W-3, then rows W-2, W-1 mirrored into W-4, W-5 */
        bs = (bits & 4) ? SEED_B : SEED_A;
        for (h = 5; h <= half; ++h) {
            float r = signed_draw(&bs);
            PUT(h, W - 3, r);
            PUT(W - 1 - h, W - 3, r);
        }
        for (v = W - 2; v < W; ++v) {
            for (h = 5; h < W - 5; ++h) {
                float r = signed_draw(&bs);
                PUT(h, v, r);
                PUT(W - 1 - h, 2 * W - 6 - v, r);
            }
        }

        /* left: column 2, then columns 1, 0 mirrored into 3, 4 */
        bs = (bits & 8) ? SEED_B : SEED_A;
        for (v = 5; v <= half; ++v) {
            float r = signed_draw(&bs);
            PUT(2, v, r);
            PUT(2, W - 1 - v, r);
        }
        for (h = 1; h >= 0; --h) {
            for (v = 5; v < W - 5; ++v) {
                float r = signed_draw(&bs);
                PUT(h, v, r);
                PUT(4 - h, W - 1 - v, r);
            }
        }

        /* centre */
        for (v = 5; v < W - 5; ++v) {
            for (h = 5; h < W - 5; ++h) {
                PUT(h, v, signed_draw(&centre_seed));
            }
        }
#undef PUT
    }
}

/* fp32 -> fp16 bits, round-to-nearest-even, with subnormals, inf and nan */
uint16_t orc_float_to_half_bits(float v)
{
    uint32_t x;
    memcpy(&x, &v, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t mant = x & 0x007FFFFFu;
    int32_t e = (int32_t) ((x >> 23) & 0xFF);

    if (e == 0xFF) {                               /* inf / nan */
        return (uint16_t) (sign | 0x7C00u | (mant ? 0x0200u : 0));
    }
    int32_t he = e - 127 + 15;
    if (he >= 0x1F) {
        return (uint16_t) (sign | 0x7C00u);        /* overflow -> inf */
    }
    if (he <= 0) {                                 /* subnormal or zero */
        if (he < -10) return (uint16_t) sign;
        mant |= 0x00800000u;
        int shift = 14 - he;                       /* 14..24 */
        uint32_t hm = mant >> shift;
        uint32_t rem = mant & ((1u << shift) - 1);
        uint32_t halfway = 1u << (shift - 1);
        if (rem > halfway || (rem == halfway && (hm & 1))) hm++;
        return (uint16_t) (sign | hm);
    }
    uint32_t hm = mant >> 13;
    uint32_t rem = mant & 0x1FFFu;
    uint32_t h = (uint32_t) (he << 10) | hm;
    if (rem > 0x1000u || (rem == 0x1000u && (hm & 1))) h++;   /* may carry into exponent: correct */
    return (uint16_t) (sign | h);
}

static float half_bits_to_float(uint16_t h)
{
    uint32_t sign = (uint32_t) (h & 0x8000u) << 16;
    uint32_t e = (h >> 10) & 0x1F;
    uint32_t m = h & 0x3FFu;
    uint32_t x;
    if (e == 0) {
        if (m == 0) {
            x = sign;
        } else {
            int sh = 0;
            while (!(m & 0x400u)) { m <<= 1; sh++; }
            m &= 0x3FFu;
            x = sign | (uint32_t) ((127 - 15 - sh + 1) << 23) | (m << 13);
        }
    } else if (e == 0x1F) {
        x = sign | 0x7F800000u | (m << 13);
    } else {
        x = sign | ((e - 15 + 127) << 23) | (m << 13);
    }
    float f;
    memcpy(&f, &x, 4);
    return f;
}

float orc_round_half(float v)
{
    return half_bits_to_float(orc_float_to_half_bits(v));
}

void orc_dem_noise_r16f(int W, float *out)
{
    orc_dem_noise(W, out);
    for (size_t i = 0; i < (size_t) 6 * W * W; ++i) out[i] = orc_round_half(out[i]);
}

/* ------------------------------------------------------------------------
 * ElevationProducer.cpp:345-373.  Four border bits from the sign of cnoise at
 * the midpoints of the tile's edges in a lattice where the six cube faces are
 * unfolded (faces 2..5 side by side, face 1 above, face 6 below); face 0 (flat
 * terrain) goes through the generic branch with offset = -2 * 2^level.
 * Integer operands are converted to float exactly as the C++ call does
 * (int + 0.5 is a double expression, narrowed to the float parameter).
 * ------------------------------------------------------------------------ */
static inline int pos(float v) { return v > 0.0f; }

void orc_noise_select(int level, int tx, int ty, int face, int *noiseR, int *noiseL)
{
    static const int rot_of[16]   = { 0, 0, 1, 0, 2, 0, 1, 0, 3, 3, 1, 3, 2, 2, 1, 0 };
    static const int layer_of[16] = { 0, 1, 1, 2, 1, 3, 2, 4, 1, 2, 3, 4, 2, 4, 4, 5 };
    int bottom, right, top, left;
    const int n = 1 << level;

    if (face == 1) {
        bottom = pos(orc_cnoise2((float) (tx + 0.5), (float) (ty + n)));
        right = (tx == n - 1)
            ? pos(orc_cnoise2((float) (ty + n + 0.5), (float) n))
            : pos(orc_cnoise2((float) (tx + 1), (float) (ty + n + 0.5)));
        top = (ty == n - 1)
            ? pos(orc_cnoise2((float) ((3 * n - 1 - tx) + 0.5), (float) n))
            : pos(orc_cnoise2((float) (tx + 0.5), (float) (ty + n + 1)));
        left = (tx == 0)
            ? pos(orc_cnoise2((float) ((4 * n - 1 - ty) + 0.5), (float) n))
            : pos(orc_cnoise2((float) tx, (float) (ty + n + 0.5)));
    } else if (face == 6) {
        bottom = (ty == 0)
            ? pos(orc_cnoise2((float) ((3 * n - 1 - tx) + 0.5), 0.0f))
            : pos(orc_cnoise2((float) (tx + 0.5), (float) (ty - n)));
        right = (tx == n - 1)
            ? pos(orc_cnoise2((float) ((2 * n - 1 - ty) + 0.5), 0.0f))
            : pos(orc_cnoise2((float) (tx + 1), (float) (ty - n + 0.5)));
        top = pos(orc_cnoise2((float) (tx + 0.5), (float) (ty - n + 1)));
        left = (tx == 0)
            ? pos(orc_cnoise2((float) (3 * n + ty + 0.5), 0.0f))
            : pos(orc_cnoise2((float) tx, (float) (ty - n + 0.5)));
    } else {
        const int off = n * (face - 2);
        bottom = pos(orc_cnoise2((float) (tx + off + 0.5), (float) ty));
        right = pos(orc_cnoise2((float) ((tx + off + 1) % (4 << level)), (float) (ty + 0.5)));
        top = pos(orc_cnoise2((float) (tx + off + 0.5), (float) (ty + 1)));
        left = pos(orc_cnoise2((float) (tx + off), (float) (ty + 0.5)));
    }
    int bits = bottom + 2 * right + 4 * top + 8 * left;
    *noiseR = rot_of[bits];
    *noiseL = layer_of[bits];
}
