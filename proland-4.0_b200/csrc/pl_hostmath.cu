/*
 * pl_hostmath.cu -- host-side integer / fp maths of the tile-production path:
 * what ElevationProducer / NormalProducer compute on the CPU before a draw.
 *
 *   31-bit LCG + frandom          core/sources/proland/math/noise.h:50-67
 *   Perlin tables                 core/sources/proland/math/noise.cpp:68-101
 *   createDemNoise                terrain/sources/proland/dem/ElevationProducer.cpp:50-128
 * plus the host entry points of the shared per-tile maths in pl_reqmath.cuh.
 *
 * Host code is compiled with -ffp-contract=off: cnoise feeds integer decisions
 * and has to round like a plain fp32 C++ build of the reference.
 */
#include <cmath>
#include <cstring>
#include <utility>

#include "pl_reqmath.cuh"

namespace {

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

struct PerlinTables {
    int perm[kPerlinN];
    float g2[kPerlinN * 2];
    PerlinTables()
    {
        Lcg rng(12345);
        auto draw = [&rng]() { return (float) ((int) (rng.next() % (2 * kPerlinB)) - kPerlinB) / kPerlinB; };
        for (int i = 0; i < kPerlinB; ++i) {
            perm[i] = i;
            (void) draw();                      /* the 1-D gradient consumes a draw */
            float gx = draw(), gy = draw();
            float len = std::sqrt(gx * gx + gy * gy);
            g2[2 * i] = gx / len;
            g2[2 * i + 1] = gy / len;
            (void) draw(); (void) draw(); (void) draw();   /* the 3-D gradient */
        }
        for (int i = kPerlinB - 1; i > 0; --i) {
            int j = (int) (rng.next() % kPerlinB);
            std::swap(perm[i], perm[j]);
        }
        for (int i = 0; i < kPerlinB + 2; ++i) {
            perm[kPerlinB + i] = perm[i];
            g2[2 * (kPerlinB + i)] = g2[2 * i];
            g2[2 * (kPerlinB + i) + 1] = g2[2 * i + 1];
        }
    }
};

const PerlinTables &tables()
{
    static const PerlinTables t;
    return t;
}

}  // namespace

PerlinView pl_host_perlin()
{
    const PerlinTables &t = tables();
    PerlinView v = { t.perm, t.g2 };
    return v;
}

extern "C" float pl_cnoise2(float x, float y) { return cnoise2(pl_host_perlin(), x, y); }

extern "C" void pl_noise_select(int level, int tx, int ty, int face, int *noiseR, int *noiseL)
{
    noise_select(pl_host_perlin(), level, tx, ty, face, noiseR, noiseL);
}

extern "C" void pl_elev_make_req(int tile_w, float root_quad_size, const float *noise_amp, int n_amp, int face,
                                 int level, int tx, int ty, int resid_tile_w, int has_resid, pl_elev_req *req)
{
    elev_fill_req(pl_host_perlin(), tile_w, root_quad_size, noise_amp, n_amp, face, level, tx, ty, resid_tile_w,
                  has_resid, req);
}

extern "C" void pl_norm_make_req(const pl_norm_scene *sc, double root_quad_size, int components, int level, int tx,
                                 int ty, pl_norm_req *req)
{
    (void) components;   /* RG8 storages never sample the parent normal tile */
    norm_fill_req(sc->sphere, root_quad_size, level, tx, ty, req);
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
            for (int h = 5; h <= mid; ++h) { float q = r.sym(); at(h, 2) = q; at(last - h, 2) = q; }
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
