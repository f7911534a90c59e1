/*
 * orc_ortho.c -- ORACLE (test infrastructure, never shipped): OrthoProducer's
 * tile production, restated on the CPU.
 *
 *   terrain/sources/proland/ortho/OrthoProducer.cpp:48-118   createOrthoNoise
 *   terrain/sources/proland/ortho/OrthoProducer.cpp:268-372  doCreateTile (uniforms)
 *   demo/shaders/ortho/upsampleOrthoShader.glsl:64-160       rgb_to_hsv, hsv_to_rgb, main
 *
 * Evaluation order: orc_fp.h (a*b+c and a-b*c are one fma, / is IEEE, everything else single
 * operations left to right as the shader spells them).
 *
 * GL behaviours assumed (Ork / GL are absent here, "parity unpinned", DESIGN.md "Oracle"):
 *   - an unorm8 texel c is fetched as the fp32 value c/255 (OpenGL 3.3 spec 2.1.2); the shader's
 *     "* 255.0" gives back exactly c for every c in 0..255 (checked exhaustively in
 *     tests/test_oracle.py), so the 4-tap upsample is integer arithmetic: floor(sum/16) = sum >> 4
 *   - all fetches hit texel centres: LINEAR and NEAREST storages read the same value
 *   - a residual texture with fewer than 4 channels reads (r, g|0, b|0, 1)
 *   - the RGBA8 colour buffer stores clamp(v, 0, 1) rounded to nearest (orc_unorm8)
 */
#include <math.h>
#include <string.h>
#include "orc.h"
#include "orc_fp.h"

/* OrthoProducer.cpp:48-118.  out = 6 layers of W*W*4 bytes.  Every layer starts as 128; the four
 * border strips (2 texels wide, between the 4-texel corners) come from their own LCG stream -- one of
 * two seeds chosen by the layer's border-pattern bit -- and are written together with their mirror
 * image; the interior [4, W-4)^2 comes from one stream that keeps running across layers.
 * int(frandom * 255.0f): values 0..254. */
void orc_ortho_noise(int W, uint8_t *out)
{
    static const int pattern[6] = { 0, 1, 3, 5, 7, 15 };
    long interior = 1234567;
    for (int nl = 0; nl < 6; ++nl) {
        uint8_t *n = out + (size_t) nl * W * W * 4;
        const int l = pattern[nl];
        long b;
        memset(n, 128, (size_t) W * W * 4);
#define PUT(h, v, c, val) n[4 * ((h) + (v) * W) + (c)] = (uint8_t) (val)
        b = (l & 1) == 0 ? 7654321 : 5647381;               /* bottom */
        for (int v = 2; v < 4; ++v)
            for (int h = 4; h < W - 4; ++h)
                for (int c = 0; c < 4; ++c) {
                    int N = (int) (orc_frandom(&b) * 255.0f);
                    PUT(h, v, c, N);
                    PUT(W - 1 - h, 3 - v, c, N);
                }
        b = (l & 2) == 0 ? 7654321 : 5647381;               /* right */
        for (int h = W - 3; h >= W - 4; --h)
            for (int v = 4; v < W - 4; ++v)
                for (int c = 0; c < 4; ++c) {
                    int N = (int) (orc_frandom(&b) * 255.0f);
                    PUT(h, v, c, N);
                    PUT(2 * W - 5 - h, W - 1 - v, c, N);
                }
        b = (l & 4) == 0 ? 7654321 : 5647381;               /* top */
        for (int v = W - 2; v < W; ++v)
            for (int h = 4; h < W - 4; ++h)
                for (int c = 0; c < 4; ++c) {
                    int N = (int) (orc_frandom(&b) * 255.0f);
                    PUT(h, v, c, N);
                    PUT(W - 1 - h, 2 * W - 5 - v, c, N);
                }
        b = (l & 8) == 0 ? 7654321 : 5647381;               /* left */
        for (int h = 1; h >= 0; --h)
            for (int v = 4; v < W - 4; ++v)
                for (int c = 0; c < 4; ++c) {
                    int N = (int) (orc_frandom(&b) * 255.0f);
                    PUT(h, v, c, N);
                    PUT(3 - h, W - 1 - v, c, N);
                }
        for (int v = 4; v < W - 4; ++v)                      /* centre */
            for (int h = 4; h < W - 4; ++h)
                for (int c = 0; c < 4; ++c) PUT(h, v, c, (int) (orc_frandom(&interior) * 255.0f));
#undef PUT
    }
}

/* OrthoProducer.cpp:286-372: the uniforms of one tile */
void orc_ortho_uniforms(int W, int face, int level, int tx, int ty, const float *noiseAmp, int nAmp,
                        const float noiseColor[4], const float rootNoiseColor[4], int hsv, float scale,
                        int hasResidual, orc_ortho_params *p)
{
    const int tileSize = W - 4;
    memset(p, 0, sizeof *p);
    p->tileWidth = W;
    p->level = level;
    p->dx = level > 0 ? (tx % 2) * (tileSize / 2) : -1;
    p->dy = level > 0 ? (ty % 2) * (tileSize / 2) : -1;
    p->hasResidual = hasResidual;
    p->residualScale = hasResidual ? scale : -1.0f;          /* residualOSH.w */
    const float rs = level < nAmp ? noiseAmp[level] : 0.0f;
    orc_noise_select(level, tx, ty, face, &p->noiseR, &p->noiseL);   /* same code as ElevationProducer's */
    p->hsv = hsv;
    if (hsv) {            /* vec4f(noiseColor) * vec4f(rs, rs, rs, scale * rs) / 255.0f */
        for (int c = 0; c < 3; ++c) p->noiseColor[c] = noiseColor[c] * rs / 255.0f;
        p->noiseColor[3] = noiseColor[3] * (scale * rs) / 255.0f;
    } else {              /* noiseColor * scale * rs / 255.0f */
        for (int c = 0; c < 4; ++c) p->noiseColor[c] = noiseColor[c] * scale * rs / 255.0f;
    }
    for (int c = 0; c < 4; ++c) p->rootNoiseColor[c] = rootNoiseColor[c];
}

static float min3(float a, float b, float c) { return fminf(a, fminf(b, c)); }
static float max3(float a, float b, float c) { return fmaxf(a, fmaxf(b, c)); }

/* upsampleOrthoShader.glsl:64-92 */
static void rgb_to_hsv(const float rgb[3], float hsv[3])
{
    hsv[0] = hsv[1] = 0.0f;
    const float minVal = min3(rgb[0], rgb[1], rgb[2]);
    const float maxVal = max3(rgb[0], rgb[1], rgb[2]);
    const float delta = maxVal - minVal;
    hsv[2] = maxVal;
    if (delta != 0.0f) {
        float del[3];
        hsv[1] = delta / maxVal;
        for (int c = 0; c < 3; ++c) del[c] = ((maxVal - rgb[c]) / 6.0f + delta / 2.0f) / delta;
        if (rgb[0] == maxVal) hsv[0] = del[2] - del[1];
        else if (rgb[1] == maxVal) hsv[0] = (float) (1.0 / 3.0) + del[0] - del[2];
        else if (rgb[2] == maxVal) hsv[0] = (float) (2.0 / 3.0) + del[1] - del[0];
        if (hsv[0] < 0.0f) hsv[0] += 1.0f;
        if (hsv[0] > 1.0f) hsv[0] -= 1.0f;
    }
}

/* upsampleOrthoShader.glsl:94-121 */
static void hsv_to_rgb(const float hsv[3], float rgb[3])
{
    rgb[0] = rgb[1] = rgb[2] = hsv[2];
    if (hsv[1] != 0.0f) {
        const float var_h = hsv[0] * 6.0f;
        const float var_i = floorf(var_h);
        const float f = var_h - var_i;
        const float var_1 = hsv[2] * (1.0f - hsv[1]);
        const float var_2 = hsv[2] * orc_fma(-hsv[1], f, 1.0f);
        const float var_3 = hsv[2] * orc_fma(-hsv[1], 1.0f - f, 1.0f);
        const float V = hsv[2];
        if (var_i == 0.0f) { rgb[0] = V; rgb[1] = var_3; rgb[2] = var_1; }
        else if (var_i == 1.0f) { rgb[0] = var_2; rgb[1] = V; rgb[2] = var_1; }
        else if (var_i == 2.0f) { rgb[0] = var_1; rgb[1] = V; rgb[2] = var_3; }
        else if (var_i == 3.0f) { rgb[0] = var_1; rgb[1] = var_2; rgb[2] = V; }
        else if (var_i == 4.0f) { rgb[0] = var_3; rgb[1] = var_1; rgb[2] = V; }
        else { rgb[0] = V; rgb[1] = var_1; rgb[2] = var_2; }
    }
}

static float clamp01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }

/* GLSL smoothstep(0.4, 0.8, x) */
static float smoothstep_04_08(float x)
{
    const float t = clamp01((x - 0.4f) / (0.8f - 0.4f));
    return t * t * orc_fma(-2.0f, t, 3.0f);
}

/* upsampleOrthoShader.glsl:123-158 for the whole tile.  parent: W*W*4 bytes (NULL at level 0),
 * residual: W*W*channels bytes (NULL: none), noise: orc_ortho_noise's 6 layers.  out: W*W*4 bytes,
 * the RGBA8 colour buffer (a storage with fewer channels keeps the first ones). */
void orc_ortho_tile(const orc_ortho_params *p, const uint8_t *parent, const uint8_t *residual, int channels,
                    const uint8_t *noise, uint8_t *out)
{
    static const float masks[4][4] = { { 1, 3, 3, 9 }, { 3, 1, 9, 3 }, { 3, 9, 1, 3 }, { 9, 3, 3, 1 } };
    const int W = p->tileWidth;
    const uint8_t *nl = noise + (size_t) p->noiseL * W * W * 4;
    for (int y = 0; y < W; ++y) {
        for (int x = 0; x < W; ++x) {
            float result[4] = { 128.0f, 128.0f, 128.0f, 128.0f };
            if (p->hasResidual && residual) {
                const uint8_t *r = residual + (size_t) (x + y * W) * channels;
                for (int c = 0; c < 4; ++c) {
                    const float texel = c < channels ? (float) r[c] / 255.0f : (c == 3 ? 1.0f : 0.0f);
                    result[c] = texel * 255.0f;
                }
            } else if (!parent) {
                for (int c = 0; c < 4; ++c) result[c] = p->rootNoiseColor[c] * 255.0f;
            }
            if (parent) {
                const float *m = masks[(x & 1) + 2 * (y & 1)];
                const int px = ((x + 1) >> 1) + p->dx, py = ((y + 1) >> 1) + p->dy;
                for (int c = 0; c < 4; ++c) {
                    const float c0 = (float) parent[4 * (px + py * W) + c] / 255.0f * 255.0f;
                    const float c1 = (float) parent[4 * (px + 1 + py * W) + c] / 255.0f * 255.0f;
                    const float c2 = (float) parent[4 * (px + (py + 1) * W) + c] / 255.0f * 255.0f;
                    const float c3 = (float) parent[4 * (px + 1 + (py + 1) * W) + c] / 255.0f * 255.0f;
                    const float s = orc_fma(m[3], c3, orc_fma(m[2], c2, orc_fma(m[1], c1, m[0] * c0)));
                    const float cc = floorf(s / 16.0f);
                    result[c] = orc_fma(result[c] - 128.0f, p->residualScale, cc);
                }
            }
            /* uvs = (nuv, 1 - nuv), NEAREST/REPEAT: texel (x, y, W-1-x, W-1-y)[R], [..][(R+1)%4] */
            const int sel[4] = { x, y, W - 1 - x, W - 1 - y };
            const uint8_t *nt = nl + 4 * (sel[p->noiseR] + sel[(p->noiseR + 1) % 4] * W);
            float n[4];
            for (int c = 0; c < 4; ++c) n[c] = (float) nt[c] / 255.0f * 255.0f;
            if (p->hsv) {
                float rgb[3], hsv[3];
                for (int c = 0; c < 3; ++c) rgb[c] = result[c] / 255.0f;
                rgb_to_hsv(rgb, hsv);
                const float k = 1.0f - smoothstep_04_08(hsv[2]);
                for (int c = 0; c < 3; ++c) hsv[c] *= 1.0f + k * p->noiseColor[c] * (n[c] - 128.0f) / 255.0f;
                hsv[0] = hsv[0] - floorf(hsv[0]);
                hsv[1] = clamp01(hsv[1]);
                hsv[2] = clamp01(hsv[2]);
                hsv_to_rgb(hsv, rgb);
                for (int c = 0; c < 3; ++c) result[c] = rgb[c] * 255.0f;
                result[3] = orc_fma(p->noiseColor[3], n[3] - 128.0f, result[3]);
            } else {
                for (int c = 0; c < 4; ++c) result[c] = orc_fma(p->noiseColor[c], n[c] - 128.0f, result[c]);
            }
            for (int c = 0; c < 4; ++c) out[4 * (x + y * W) + c] = orc_unorm8(result[c] / 255.0f);
        }
    }
}

/* A whole quadtree of ortho tiles, levels 0..maxLevel, breadth first (no residuals): out holds the
 * tiles in level order, Morton order inside a level (x in the even bits).  Returns the tile count.
 * Used by bench-style CPU timings and the parity tests. */
static unsigned demorton(uint64_t m, int odd)
{
    unsigned v = 0;
    for (int b = 0; b < 24; ++b) v |= (unsigned) ((m >> (2 * b + odd)) & 1) << b;
    return v;
}

long orc_ortho_quadtree(int W, int face, int maxLevel, const float *noiseAmp, int nAmp, const float noiseColor[4],
                        const float rootNoiseColor[4], int hsv, float scale, const uint8_t *noise, uint8_t *out)
{
    const size_t tb = (size_t) W * W * 4;
    long done = 0;
    size_t levelStart = 0, parentStart = 0;
    for (int level = 0; level <= maxLevel; ++level) {
        const long n = 1L << (2 * level);
#pragma omp parallel for schedule(static)
        for (long m = 0; m < n; ++m) {
            const int tx = (int) demorton((uint64_t) m, 0), ty = (int) demorton((uint64_t) m, 1);
            orc_ortho_params p;
            orc_ortho_uniforms(W, face, level, tx, ty, noiseAmp, nAmp, noiseColor, rootNoiseColor, hsv, scale, 0, &p);
            const uint8_t *parent = level > 0 ? out + (parentStart + (size_t) (m >> 2)) * tb : NULL;
            orc_ortho_tile(&p, parent, NULL, 4, noise, out + (levelStart + (size_t) m) * tb);
        }
        done += n;
        parentStart = levelStart;
        levelStart += (size_t) n;
    }
    return done;
}

/* ------------------------------------------------------------------------------------------------
 * OrthoCPUProducer (ortho/OrthoCPUProducer.cpp:68-118, 160-246): the reader of ortho residual files.
 *   header  : 7 int32 (maxLevel, tileSize, channels, rootLevel, rootTx, rootTy, flags); flags & 1 = DXT
 *             blobs, flags & 2 = no border
 *   offsets : ntiles x (begin, end) int64 relative to the end of the table, ntiles = (4^(maxLevel+1) - 1) / 3
 *   blob    : tile id = tx + ty * 2^level + (4^level - 1) / 3; a TIFF whose single strip inflates to
 *             (tileSize + 2 border)^2 * channels bytes (TIFFReadEncodedStrip)
 * Returns the tile width, or < 0: -1 bad header / level, -2 DXT file (not decoded on the CPU either: the
 * reference hands those bytes to the GL texture unit), -3 corrupt blob. */
int orc_ortho_cpu_read(const uint8_t *file, size_t size, int level, int tx, int ty, uint8_t *out, int *channels)
{
    if (size < 28) return -1;
    int32_t h[7];
    memcpy(h, file, 28);
    const int maxLevel = h[0], tileSize = h[1], ch = h[2], flags = h[6];
    if (level < 0 || level > maxLevel || ch < 1 || ch > 4) return -1;
    if (flags & 1) return -2;
    const int border = (flags & 2) ? 0 : 2;
    const long ntiles = ((1L << (maxLevel * 2 + 2)) - 1) / 3;
    const size_t header = 28 + (size_t) ntiles * 16;
    if (size < header) return -1;
    const int tileid = tx + ty * (1 << level) + ((1 << (2 * level)) - 1) / 3;
    int64_t range[2];
    memcpy(range, file + 28 + (size_t) tileid * 16, 16);
    if (range[0] < 0 || range[1] < range[0] || header + (size_t) range[1] > size) return -3;
    const int w = tileSize + 2 * border;
    int tw = 0, th = 0;
    const long got = orc_tiff_inflate(file + header + range[0], (uint32_t) (range[1] - range[0]), out,
                                      (size_t) w * w * ch, &tw, &th);
    if (got != (long) w * w * ch || tw != w || th != w) return -3;
    if (channels) *channels = ch;
    return w;
}
