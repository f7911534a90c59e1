/*
 * orc_normal.c -- ORACLE (test infrastructure, never shipped): the
 * NormalProducer pass.
 *
 * Restates:
 *   terrain/sources/proland/dem/NormalProducer.cpp:175-283 (uniforms; the
 *       spherical patch geometry is evaluated in double and narrowed to fp32)
 *   src/demo/shaders/elevation/normalShader.glsl:60-125    (flat + sphere,
 *       four output formats, optional parent coarse normal); the example
 *       variants N-flat / N-sphere are the format-3 special cases.
 *
 * Sampler semantics (OpenGL 3.3 spec 3.8.8): the shader fetches the elevation
 * tile at texel coordinate (i + 0.25); with a NEAREST storage that is texel i,
 * with a LINEAR storage it is 0.25*T[i-1] + 0.75*T[i] on each axis.
 * Ork (absent) supplies vec3d::normalize; restated as v * (1/|v|).
 * fp32 shader maths follows the canonical order of orc_fp.h (fma chains for
 * dot / matrix*vector / a*b+c; -DORC_STRICT: no contraction).
 */
#include "orc.h"
#include "orc_fp.h"
#include <math.h>
#include <string.h>

typedef struct { double x, y, z; } d3;

static d3 d3_normalize(d3 v, double *len)
{
    double l = sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
    double inv = 1.0 / l;
    d3 r = { v.x * inv, v.y * inv, v.z * inv };
    if (len) *len = l;
    return r;
}
static d3 d3_cross(d3 a, d3 b)
{
    d3 r = { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x };
    return r;
}

/* tangent frame at the point (px,py,R) of the cube face: rows ux, uy, uz
 * (NormalProducer.cpp:243-249; core/doc/overview.txt:886-896) */
static void tangent_rows(double px, double py, double R, d3 *ux, d3 *uy, d3 *uz)
{
    d3 pc = { px, py, R };
    d3 unit_y = { 0.0, 1.0, 0.0 };
    *uz = d3_normalize(pc, NULL);
    *ux = d3_normalize(d3_cross(unit_y, *uz), NULL);
    *uy = d3_cross(*uz, *ux);
}

void orc_normal_uniforms(int W, int gridMeshSize, int components, int signed_comp,
                         int elev_W, int elev_border, int elev_filter, int parent_filter,
                         double rootQuadSize, int sphere,
                         int level, int tx, int ty, orc_norm_params *p)
{
    memset(p, 0, sizeof(*p));
    p->W = W;
    p->grid = (W - 1) / gridMeshSize;
    p->format = components == 4 ? (signed_comp ? 0 : 1) : (signed_comp ? 2 : 3);
    p->elev_W = elev_W;
    p->elev_border = elev_border;
    p->elev_filter = elev_filter;
    p->has_parent = (level > 0 && components == 4);
    p->ptx = tx % 2;
    p->pty = ty % 2;
    p->parent_filter = parent_filter;

    const double D = rootQuadSize;
    const double R = D / 2.0;
    const double n = (double) (1 << level);
    const double x0 = (double) tx / n * D - R;
    const double y0 = (double) ty / n * D - R;

    if (!sphere) {
        static const float ident[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
        memcpy(p->w2t, ident, sizeof(ident));
        memcpy(p->p2t, ident, sizeof(ident));
        p->deform[0] = (float) x0;
        p->deform[1] = (float) y0;
        p->deform[2] = (float) (D / n);
        p->deform[3] = 0.0f;
        return;
    }

    const double x1 = (double) (tx + 1) / n * D - R;
    const double y1 = (double) (ty + 1) / n * D - R;
    d3 c[4] = { { x0, y0, R }, { x1, y0, R }, { x0, y1, R }, { x1, y1, R } };
    d3 v[4];
    double l[4];
    for (int k = 0; k < 4; ++k) v[k] = d3_normalize(c[k], &l[k]);
    d3 vc = { (v[0].x + v[1].x + v[2].x + v[3].x) * 0.25,
              (v[0].y + v[1].y + v[2].y + v[3].y) * 0.25,
              (v[0].z + v[1].z + v[2].z + v[3].z) * 0.25 };

    for (int k = 0; k < 4; ++k) {
        /* rows: x, y, z, 1 ; columns: the four corners */
        p->corners[0 * 4 + k] = (float) (v[k].x * R - vc.x * R);
        p->corners[1 * 4 + k] = (float) (v[k].y * R - vc.y * R);
        p->corners[2 * 4 + k] = (float) (v[k].z * R - vc.z * R);
        p->corners[3 * 4 + k] = 1.0f;
        p->verticals[0 * 4 + k] = (float) v[k].x;
        p->verticals[1 * 4 + k] = (float) v[k].y;
        p->verticals[2 * 4 + k] = (float) v[k].z;
        p->verticals[3 * 4 + k] = 0.0f;
        p->norms[k] = (float) l[k];
    }

    d3 ux, uy, uz;
    tangent_rows((x0 + x1) * 0.5, (y0 + y1) * 0.5, R, &ux, &uy, &uz);
    double w2t[9] = { ux.x, ux.y, ux.z, uy.x, uy.y, uy.z, uz.x, uz.y, uz.z };
    for (int k = 0; k < 9; ++k) p->w2t[k] = (float) w2t[k];

    if (level > 0) {
        const double np = (double) (1 << (level - 1));
        const double px0 = (tx / 2 + 0.5) / np * D - R;
        const double py0 = (ty / 2 + 0.5) / np * D - R;
        d3 pux, puy, puz;
        tangent_rows(px0, py0, R, &pux, &puy, &puz);
        /* parent tangent -> world: columns are the parent's ux, uy, uz */
        double t2w[9] = { pux.x, puy.x, puz.x, pux.y, puy.y, puz.y, pux.z, puy.z, puz.z };
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) {
                double s = w2t[i * 3 + 0] * t2w[0 * 3 + j] + w2t[i * 3 + 1] * t2w[1 * 3 + j]
                         + w2t[i * 3 + 2] * t2w[2 * 3 + j];
                p->p2t[i * 3 + j] = (float) s;
            }
        }
    }

    p->deform[0] = (float) x0;
    p->deform[1] = (float) y0;
    p->deform[2] = (float) (D / n);
    p->deform[3] = (float) R;
}

/* ---------------------------------------------------------------------- */

static inline int clampi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }
static inline int floordiv(int a, int b) { int q = a / b; return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q; }

/* elevation zm fetch at texel coordinate (i + 0.25, j + 0.25) */
static float fetch_zm(const orc_norm_params *p, const float *elev, int i, int j)
{
    const int EW = p->elev_W;
#define ZM(a, b) elev[(clampi(a, 0, EW - 1) + clampi(b, 0, EW - 1) * EW) * 3 + 2]
    if (p->elev_filter == ORC_FILTER_NEAREST) {
        return ZM(i, j);
    }
    /* GL_LINEAR: u - 0.5 = i - 0.25 -> i0 = i-1, frac = 0.75 on both axes;
     * tau = (1-a)(1-b) t00 + a(1-b) t10 + (1-a) b t01 + a b t11 (spec eq. 3.26) */
    const float a = 0.75f, b = 0.75f;
    float t00 = ZM(i - 1, j - 1), t10 = ZM(i, j - 1), t01 = ZM(i - 1, j), t11 = ZM(i, j);
    return orc_fma(a * b, t11, orc_fma((1.0f - a) * b, t01,
           orc_fma(a * (1.0f - b), t10, ((1.0f - a) * (1.0f - b)) * t00)));
#undef ZM
}

/* parent normal fetch (.xy) at texel coordinate (i + off + 0.25) */
static void fetch_parent_xy(const orc_norm_params *p, const float *parent, float cx, float cy, float *o)
{
    const int W = p->W;
#define PN(a, b, ch) parent[(clampi(a, 0, W - 1) + clampi(b, 0, W - 1) * W) * 4 + (ch)]
    if (p->parent_filter == ORC_FILTER_NEAREST) {
        int i = (int) floorf(cx), j = (int) floorf(cy);
        o[0] = PN(i, j, 0);
        o[1] = PN(i, j, 1);
        return;
    }
    float fx = cx - 0.5f, fy = cy - 0.5f;
    int i0 = (int) floorf(fx), j0 = (int) floorf(fy);
    float a = fx - (float) i0, b = fy - (float) j0;
    for (int ch = 0; ch < 2; ++ch) {
        o[ch] = orc_fma(a * b, PN(i0 + 1, j0 + 1, ch), orc_fma((1.0f - a) * b, PN(i0, j0 + 1, ch),
                orc_fma(a * (1.0f - b), PN(i0 + 1, j0, ch), ((1.0f - a) * (1.0f - b)) * PN(i0, j0, ch))));
    }
#undef PN
}

/* mat4 (row-major maths matrix) times vec4, summed column by column as GLSL's
 * M * v = M[0]*v.x + M[1]*v.y + M[2]*v.z + M[3]*v.w */
static inline void m4v(const float *m, const float *v, float *o)
{
    for (int r = 0; r < 4; ++r) {
        o[r] = orc_dot4(m + r * 4, v);
    }
}
static inline void m3v(const float *m, const float *v, float *o)
{
    for (int r = 0; r < 3; ++r) {
        o[r] = orc_dot3(m + r * 3, v);
    }
}

/* normalShader.glsl:60-82 */
static void world_position(const orc_norm_params *p, float ux, float uy, float h, float *pos)
{
    float u = ux / ((float) p->W - 1.0f);
    float v = uy / ((float) p->W - 1.0f);
    if (p->deform[3] == 0.0f) {
        pos[0] = orc_fma(p->deform[2], u, p->deform[0]);
        pos[1] = orc_fma(p->deform[2], v, p->deform[1]);
        pos[2] = h;
        return;
    }
    const float R = p->deform[3];
    const float *L = p->norms;
    float U = 1.0f - u, V = 1.0f - v;
    float alpha[4] = { U * V, u * V, U * v, u * v };       /* uvUV.zxzx * uvUV.wwyy */
    float al[4] = { alpha[0] * L[0], alpha[1] * L[1], alpha[2] * L[2], alpha[3] * L[3] };
    float den = orc_dot4(alpha, L);
    float ap[4] = { al[0] / den, al[1] / den, al[2] / den, al[3] / den };

    float up[4], base[4];
    m4v(p->verticals, ap, up);
    /* smoothstep(R/32, R/64, deform.z) */
    float e0 = R / 32.0f, e1 = R / 64.0f;
    float t = (p->deform[2] - e0) / (e1 - e0);
    t = fminf(fmaxf(t, 0.0f), 1.0f);
    float s = t * t * orc_fma(-2.0f, t, 3.0f);
    float len = sqrtf(orc_dot3(up, up));
    float k = orc_fma(1.0f, s, len * (1.0f - s));           /* mix(len, 1, s) */
    float hPrime = orc_fma(R, 1.0f - k, h) / k;
    m4v(p->corners, ap, base);
    pos[0] = orc_fma(hPrime, up[0], base[0]);
    pos[1] = orc_fma(hPrime, up[1], base[1]);
    pos[2] = orc_fma(hPrime, up[2], base[2]);
}

/* normalShader.glsl:84-125 */
void orc_normal_tile(const orc_norm_params *p, const float *elev,
                     const float *parent, float *out)
{
    const int W = p->W;
    const int b = p->elev_border;
    const int g = p->grid;

    for (int y = 0; y < W; ++y) {
        for (int x = 0; x < W; ++x) {
            float z0 = fetch_zm(p, elev, x - 1 + b, y + b);
            float z1 = fetch_zm(p, elev, x + 1 + b, y + b);
            float z2 = fetch_zm(p, elev, x + b, y - 1 + b);
            float z3 = fetch_zm(p, elev, x + b, y + 1 + b);

            float p0[3], p1[3], p2[3], p3[3];
            world_position(p, (float) x - 1.0f, (float) y, z0, p0);
            world_position(p, (float) x + 1.0f, (float) y, z1, p1);
            world_position(p, (float) x, (float) y - 1.0f, z2, p2);
            world_position(p, (float) x, (float) y + 1.0f, z3, p3);

            float a[3] = { p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2] };
            float c[3] = { p3[0] - p2[0], p3[1] - p2[1], p3[2] - p2[2] };
            float n[3] = { orc_fma(a[1], c[2], -(a[2] * c[1])), orc_fma(a[2], c[0], -(a[0] * c[2])),
                           orc_fma(a[0], c[1], -(a[1] * c[0])) };
            float inv = 1.0f / sqrtf(orc_dot3(n, n));
            n[0] *= inv; n[1] *= inv; n[2] *= inv;
            float nt[3];
            m3v(p->w2t, n, nt);
            float nf[2] = { nt[0], nt[1] };

            float nc[2] = { nf[0], nf[1] };
            if (p->has_parent && parent != NULL) {
                /* uvc = g * floor(p_uv/(2g) + (.5, 0, 0, .5)); texel coord adds
                 * (t%2) * W/2.0 + 0.25 (NormalProducer.cpp:201-205) */
                float offx = (float) p->ptx * ((float) W / 2.0f) + 0.25f;
                float offy = (float) p->pty * ((float) W / 2.0f) + 0.25f;
                float ax = (float) (g * floordiv(x + g, 2 * g));
                float ay = (float) (g * floordiv(y, 2 * g));
                float bx = (float) (g * floordiv(x, 2 * g));
                float by = (float) (g * floordiv(y + g, 2 * g));
                float nc0[2], nc1[2];
                fetch_parent_xy(p, parent, ax + offx, ay + offy, nc0);
                fetch_parent_xy(p, parent, bx + offx, by + offy, nc1);
                nc[0] = (nc0[0] + nc1[0]) * 0.5f;
                nc[1] = (nc0[1] + nc1[1]) * 0.5f;
                if (p->format == 1) {
                    nc[0] = orc_fma(nc[0], 2.0f, -1.0f);
                    nc[1] = orc_fma(nc[1], 2.0f, -1.0f);
                }
                if (p->deform[3] != 0.0f) {
                    float v3[3] = { nc[0], nc[1], sqrtf(1.0f - orc_dot2(nc, nc)) };
                    float r3[3];
                    m3v(p->p2t, v3, r3);
                    nc[0] = r3[0];
                    nc[1] = r3[1];
                }
            }

            float *o = out + (size_t) (x + y * W) * 4;
            switch (p->format) {
            case 0: o[0] = nf[0]; o[1] = nf[1]; o[2] = nc[0]; o[3] = nc[1]; break;
            case 1: o[0] = orc_fma(nf[0], 0.5f, 0.5f); o[1] = orc_fma(nf[1], 0.5f, 0.5f);
                    o[2] = orc_fma(nc[0], 0.5f, 0.5f); o[3] = orc_fma(nc[1], 0.5f, 0.5f); break;
            case 2: o[0] = nf[0]; o[1] = nf[1]; o[2] = 0.0f; o[3] = 0.0f; break;
            default: o[0] = orc_fma(nf[0], 0.5f, 0.5f); o[1] = orc_fma(nf[1], 0.5f, 0.5f); o[2] = 0.5f; o[3] = 0.5f; break;
            }
        }
    }
}

uint8_t orc_unorm8(float f)
{
    if (!(f > 0.0f)) return 0;          /* also NaN -> 0 */
    if (f >= 1.0f) return 255;
    return (uint8_t) lrintf(f * 255.0f); /* round to nearest (even on ties) */
}

void orc_pack_unorm8(int W, int channels, const float *data, uint8_t *out)
{
    for (int i = 0; i < W * W; ++i) {
        for (int c = 0; c < channels; ++c) out[i * channels + c] = orc_unorm8(data[i * 4 + c]);
    }
}
