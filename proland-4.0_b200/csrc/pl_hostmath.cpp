/*
 * pl_hostmath.cpp -- host-side integer / fp maths of the tile-production path:
 * what ElevationProducer / NormalProducer compute on the CPU before a draw.
 *
 *   31-bit LCG + frandom          core/sources/proland/math/noise.h:50-67
 *   Perlin tables + 2-D cnoise    core/sources/proland/math/noise.cpp:68-165
 *   createDemNoise                terrain/sources/proland/dem/ElevationProducer.cpp:50-128
 *   noise layer / rotation select terrain/sources/proland/dem/ElevationProducer.cpp:345-373
 *   per-tile elevation uniforms   terrain/sources/proland/dem/ElevationProducer.cpp:305-343
 *   per-tile normal uniforms      terrain/sources/proland/dem/NormalProducer.cpp:196-283
 *
 * Built with -ffp-contract=off: cnoise feeds integer decisions and has to round
 * like a plain fp32 C++ build of the reference.
 */
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

#include "pl_internal.h"

namespace {

/* ---- LCG -------------------------------------------------------------- */
struct Lcg {
    uint32_t s;
    explicit Lcg(uint32_t seed) : s(seed) {}
    uint32_t next()
    {
        /* 31-bit state; the product is taken in 64 bits and masked, which equals
         * the 32-bit wrap-around of a Win32 `long` */
        s = (uint32_t) (((uint64_t) s * 1103515245ull + 12345ull) & 0x7FFFFFFFull);
        return s;
    }
    float unit() { return (float) (next() >> 7) / 16777216.0f; }   /* [0,1) */
    float sym() { return unit() * 2.0f - 1.0f; }                    /* [-1,1) */
};

/* ---- classic Perlin lattice -------------------------------------------- */
constexpr int kB = 256;

struct PerlinTables {
    int perm[2 * kB + 2];
    float g2[2 * kB + 2][2];
    PerlinTables()
    {
        Lcg rng(12345);
        auto draw = [&rng]() { return (float) ((int) (rng.next() % (2 * kB)) - kB) / kB; };
        for (int i = 0; i < kB; ++i) {
            perm[i] = i;
            (void) draw();                      /* the 1-D gradient consumes a draw */
            float gx = draw(), gy = draw();
            float len = std::sqrt(gx * gx + gy * gy);
            g2[i][0] = gx / len;
            g2[i][1] = gy / len;
            (void) draw(); (void) draw(); (void) draw();   /* the 3-D gradient */
        }
        for (int i = kB - 1; i > 0; --i) {
            int j = (int) (rng.next() % kB);
            std::swap(perm[i], perm[j]);
        }
        for (int i = 0; i < kB + 2; ++i) {
            perm[kB + i] = perm[i];
            g2[kB + i][0] = g2[i][0];
            g2[kB + i][1] = g2[i][1];
        }
    }
};

const PerlinTables &tables()
{
    static const PerlinTables t;
    return t;
}

inline float fade(float t) { return t * t * (3.0f - 2.0f * t); }
inline float mixf(float t, float a, float b) { return a + t * (b - a); }

struct Axis { int b0, b1; float r0, r1; };
inline Axis split(float v)
{
    float t = v + 4096.0f;
    Axis a;
    a.b0 = ((int) t) & (kB - 1);                 /* truncation, like the reference */
    a.b1 = (a.b0 + 1) & (kB - 1);
    a.r0 = t - (float) (int) std::floor(t);      /* ... but floor for the fraction */
    a.r1 = a.r0 - 1.0f;
    return a;
}

inline int positive(float v) { return v > 0.0f ? 1 : 0; }

}  // namespace

extern "C" float pl_cnoise2(float x, float y)
{
    const PerlinTables &T = tables();
    const Axis ax = split(x), ay = split(y);
    const int i = T.perm[ax.b0], j = T.perm[ax.b1];
    const int b00 = T.perm[i + ay.b0], b10 = T.perm[j + ay.b0];
    const int b01 = T.perm[i + ay.b1], b11 = T.perm[j + ay.b1];
    const float sx = fade(ax.r0), sy = fade(ay.r0);

    float u = ax.r0 * T.g2[b00][0] + ay.r0 * T.g2[b00][1];
    float v = ax.r1 * T.g2[b10][0] + ay.r0 * T.g2[b10][1];
    const float lo = mixf(sx, u, v);
    u = ax.r0 * T.g2[b01][0] + ay.r1 * T.g2[b01][1];
    v = ax.r1 * T.g2[b11][0] + ay.r1 * T.g2[b11][1];
    const float hi = mixf(sx, u, v);
    return mixf(sy, lo, hi);
}

/* exported for the device-side request generator: the lattice tables */
extern "C" void pl_perlin_tables(int *perm514, float *g2_514x2)
{
    const PerlinTables &T = tables();
    std::memcpy(perm514, T.perm, sizeof(T.perm));
    std::memcpy(g2_514x2, T.g2, sizeof(T.g2));
}

extern "C" void pl_noise_select(int level, int tx, int ty, int face, int *noiseR, int *noiseL)
{
    /* bit0 bottom, bit1 right, bit2 top, bit3 left; the cube faces are unfolded
     * into one integer lattice so that the shared edge of two tiles gets the
     * same bit on both sides */
    static const int kRot[16] = { 0, 0, 1, 0, 2, 0, 1, 0, 3, 3, 1, 3, 2, 2, 1, 0 };
    static const int kLayer[16] = { 0, 1, 1, 2, 1, 3, 2, 4, 1, 2, 3, 4, 2, 4, 4, 5 };
    const int n = 1 << level;
    auto N = [](double a, double b) { return positive(pl_cnoise2((float) a, (float) b)); };
    int bottom, right, top, left;
    if (face == 1) {
        bottom = N(tx + 0.5, ty + n);
        right = tx == n - 1 ? N(ty + n + 0.5, n) : N(tx + 1, ty + n + 0.5);
        top = ty == n - 1 ? N((3 * n - 1 - tx) + 0.5, n) : N(tx + 0.5, ty + n + 1);
        left = tx == 0 ? N((4 * n - 1 - ty) + 0.5, n) : N(tx, ty + n + 0.5);
    } else if (face == 6) {
        bottom = ty == 0 ? N((3 * n - 1 - tx) + 0.5, 0) : N(tx + 0.5, ty - n);
        right = tx == n - 1 ? N((2 * n - 1 - ty) + 0.5, 0) : N(tx + 1, ty - n + 0.5);
        top = N(tx + 0.5, ty - n + 1);
        left = tx == 0 ? N(3 * n + ty + 0.5, 0) : N(tx, ty - n + 0.5);
    } else {
        const int off = n * (face - 2);
        bottom = N(tx + off + 0.5, ty);
        right = N((tx + off + 1) % (4 << level), ty + 0.5);   /* C remainder, may be < 0 */
        top = N(tx + off + 0.5, ty + 1);
        left = N(tx + off, ty + 0.5);
    }
    const int bits = bottom | (right << 1) | (top << 2) | (left << 3);
    *noiseR = kRot[bits];
    *noiseL = kLayer[bits];
}

/* Six W x W layers.  Each of the four borders of a layer is drawn from its own
 * LCG stream (one of two seeds, chosen by the layer's border-pattern bit) and
 * written together with its mirror image so that two tiles sharing an edge see
 * the same values; the interior comes from a single stream that keeps running
 * from layer to layer.  Corner 5x5 blocks stay zero. */
void pl_host_dem_noise(int W, float *out6)
{
    static const int kPattern[6] = { 0, 1, 3, 5, 7, 15 };
    const uint32_t seedFor[2] = { 7654321u, 5647381u };
    const int last = W - 1, mid = W / 2;
    Lcg interior(1234567u);
    for (int layer = 0; layer < 6; ++layer) {
        float *n = out6 + (size_t) layer * W * W;
        std::memset(n, 0, sizeof(float) * W * W);
        auto at = [n, W](int x, int y) -> float & { return n[x + y * W]; };
        const int bits = kPattern[layer];

        {   /* bottom edge: rows 0..4 */
            Lcg r(seedFor[bits & 1]);
            for (int h = 5; h <= mid; ++h) { float v = r.sym(); at(h, 2) = v; at(last - h, 2) = v; }
            for (int v = 3; v < 5; ++v)
                for (int h = 5; h < W - 5; ++h) { float q = r.sym(); at(h, v) = q; at(last - h, 4 - v) = q; }
        }
        {   /* right edge: columns W-5..W-1 */
            Lcg r(seedFor[(bits >> 1) & 1]);
            for (int v = 5; v <= mid; ++v) { float q = r.sym(); at(W - 3, v) = q; at(W - 3, last - v) = q; }
            for (int h = W - 4; h >= W - 5; --h)
                for (int v = 5; v < W - 5; ++v) { float q = r.sym(); at(h, v) = q; at(2 * W - 6 - h, last - v) = q; }
        }
        {   /* top edge: rows W-5..W-1 */
            Lcg r(seedFor[(bits >> 2) & 1]);
            for (int h = 5; h <= mid; ++h) { float q = r.sym(); at(h, W - 3) = q; at(last - h, W - 3) = q; }
            for (int v = W - 2; v < W; ++v)
                for (int h = 5; h < W - 5; ++h) { float q = r.sym(); at(h, v) = q; at(last - h, 2 * W - 6 - v) = q; }
        }
        {   /* left edge: columns 0..4 */
            Lcg r(seedFor[(bits >> 3) & 1]);
            for (int v = 5; v <= mid; ++v) { float q = r.sym(); at(2, v) = q; at(2, last - v) = q; }
            for (int h = 1; h >= 0; --h)
                for (int v = 5; v < W - 5; ++v) { float q = r.sym(); at(h, v) = q; at(4 - h, last - v) = q; }
        }
        for (int v = 5; v < W - 5; ++v)
            for (int h = 5; h < W - 5; ++h) at(h, v) = interior.sym();
    }
}

extern "C" void pl_elev_make_req(int tile_w, float root_quad_size, const float *noise_amp, int n_amp,
                                 int face, int level, int tx, int ty, int resid_tile_w, int has_resid,
                                 pl_elev_req *req)
{
    const int tileSize = tile_w - 5;
    std::memset(req, 0, sizeof(*req));
    req->out_slot = -1;
    req->parent_slot = -1;
    req->resid_slot = -1;
    req->dx = (tx % 2) * (tileSize / 2);
    req->dy = (ty % 2) * (tileSize / 2);
    if (has_resid && resid_tile_w > 0) {
        const int mod = (resid_tile_w - 5) / tileSize;
        req->rx = (tx % mod) * tileSize;
        req->ry = (ty % mod) * tileSize;
    }
    req->rs = level < n_amp ? noise_amp[level] : 0.0f;
    /* float / int / int, as getRootQuadSize() / (1 << level) / tileSize evaluates */
    req->pixel_size = root_quad_size / (float) (1 << level) / (float) tileSize;
    pl_noise_select(level, tx, ty, face, &req->noise_r, &req->noise_l);
    req->level = level;
    req->tx = tx;
    req->ty = ty;
}

namespace {
struct V3 { double x, y, z; };
inline V3 unit(V3 v, double *len = nullptr)
{
    double l = std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
    double inv = 1.0 / l;
    if (len) *len = l;
    return { v.x * inv, v.y * inv, v.z * inv };
}
inline V3 cross(V3 a, V3 b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
/* rows of the world -> tangent frame at the cube-face point (px, py, R) */
inline void frame(double px, double py, double R, V3 &ux, V3 &uy, V3 &uz)
{
    uz = unit({ px, py, R });
    ux = unit(cross({ 0.0, 1.0, 0.0 }, uz));
    uy = cross(uz, ux);
}
}  // namespace

extern "C" void pl_norm_make_req(const pl_norm_scene *sc, double root_quad_size, int components,
                                 int level, int tx, int ty, pl_norm_req *req)
{
    std::memset(req, 0, sizeof(*req));
    req->out_slot = req->elev_slot = -1;
    req->parent_slot = -1;
    req->ptx = tx % 2;
    req->pty = ty % 2;
    req->level = level;
    (void) components;

    const double D = root_quad_size, R = D / 2.0;
    const double n = (double) (1 << level);
    const double x0 = (double) tx / n * D - R, y0 = (double) ty / n * D - R;
    req->deform[0] = (float) x0;
    req->deform[1] = (float) y0;
    req->deform[2] = (float) (D / n);
    const float ident[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
    std::memcpy(req->w2t, ident, sizeof(ident));
    std::memcpy(req->p2t, ident, sizeof(ident));
    if (!sc->sphere) {
        req->deform[3] = 0.0f;
        return;
    }
    req->deform[3] = (float) R;

    const double x1 = (double) (tx + 1) / n * D - R, y1 = (double) (ty + 1) / n * D - R;
    const V3 corner[4] = { { x0, y0, R }, { x1, y0, R }, { x0, y1, R }, { x1, y1, R } };
    V3 v[4];
    double len[4];
    for (int k = 0; k < 4; ++k) v[k] = unit(corner[k], &len[k]);
    const V3 vc = { (v[0].x + v[1].x + v[2].x + v[3].x) * 0.25, (v[0].y + v[1].y + v[2].y + v[3].y) * 0.25,
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
    V3 ux, uy, uz;
    frame((x0 + x1) * 0.5, (y0 + y1) * 0.5, R, ux, uy, uz);
    const double w2t[9] = { ux.x, ux.y, ux.z, uy.x, uy.y, uy.z, uz.x, uz.y, uz.z };
    for (int k = 0; k < 9; ++k) req->w2t[k] = (float) w2t[k];
    if (level > 0) {
        const double np = (double) (1 << (level - 1));
        V3 pux, puy, puz;
        frame((tx / 2 + 0.5) / np * D - R, (ty / 2 + 0.5) / np * D - R, R, pux, puy, puz);
        const double t2w[9] = { pux.x, puy.x, puz.x, pux.y, puy.y, puz.y, pux.z, puy.z, puz.z };
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                req->p2t[i * 3 + j] = (float) (w2t[i * 3 + 0] * t2w[0 * 3 + j] + w2t[i * 3 + 1] * t2w[1 * 3 + j]
                                               + w2t[i * 3 + 2] * t2w[2 * 3 + j]);
    }
    /* smoothstep(R/32, R/64, deform.z) is tile-uniform: evaluate it once here in
     * fp32, in the canonical order the kernel and oracle use */
    const float Rf = req->deform[3];
    const float e0 = Rf / 32.0f, e1 = Rf / 64.0f;
    float t = (req->deform[2] - e0) / (e1 - e0);
    t = std::fmin(std::fmax(t, 0.0f), 1.0f);
    req->smooth = t * t * std::fma(-2.0f, t, 3.0f);
}
