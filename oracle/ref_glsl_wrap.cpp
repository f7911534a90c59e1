/*
 * ref_glsl_wrap.cpp -- ORACLE support (test infrastructure, never shipped).
 *
 * The reference's own fragment shaders of the tile-production path, compiled
 * UNCHANGED as C++ behind ref_shim/glsl_shim.h into oracle/_ref/libref_glsl.so
 * (oracle/Makefile, target `ref`; the shader files are read where they lie in
 * the reference checkout, the only edit is ref_shim/glsl_to_cxx.sed rewriting
 * the array-constructor syntax `T[4] ( ... );` to `{ ... };`).  Each variant
 * is included into a namespace of its own; its uniforms / `in` / `out`
 * variables are that namespace's (thread_local) variables.
 *
 * This file sets the uniforms the way the reference's producers do and runs
 * `main` once per fragment (st = fragment centre, what the pass-through vertex
 * shader interpolates):
 *   ElevationProducer.cpp:305-343,376   NormalProducer.cpp:195-283
 *   OrthoProducer.cpp:298-372
 * It is what pins the float part of oracle/orc_*.c: tests/test_oracle.py
 * compares the restatement with it texel by texel.
 */
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include "../orc.h"
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "glsl_shim.h"

namespace glsl {
namespace up_a { /* terrain1: plain noise */
#include "upsample_a.inc"
}
namespace up_b { /* terrain2 / preprocess: slope + curvature noise */
#include "upsample_b.inc"
}
namespace up_c { /* terrain4: plain noise, flipped diagonals */
#include "upsample_c.inc"
}
namespace up_d { /* demo: slope + curvature noise, flipped diagonals, clamp */
#include "upsample_d.inc"
}
namespace up_d_noclamp {
#define NO_CLAMP
#include "upsample_d.inc"
#undef NO_CLAMP
}
namespace nrm_flat { /* terrain1 */
#include "normal_flat.inc"
}
namespace nrm_sphere { /* terrain2 / preprocess */
#include "normal_sphere.inc"
}
namespace nrm_demo { /* demo: flat + sphere, four formats, parent coarse normal */
#include "normal_demo.inc"
}
namespace ortho_demo {
#include "ortho_demo.inc"
}
namespace ortho_ex { /* the examples' copy (noiseUVLH.w == 1) */
#include "ortho_ex.inc"
}
} /* namespace glsl */

#undef uniform
#undef in
#undef out
#undef main
#undef layout

using namespace glsl;

namespace {

struct UpArgs {
    const orc_elev_params *p;
    const float *parent;     /* W*W*3 or NULL */
    int parent_filter;
    const float *resid;      /* the residual tile (stride p->resid_stride) or NULL */
    const float *noise;      /* 6*W*W, fp16-rounded */
    int subtexel_bits;
    float *out;              /* W*W*3 */
};

/* ElevationProducer.cpp:305-343,376 + one shader invocation per fragment */
#define UPSAMPLE_RUNNER(NS)                                                                              \
    void run_##NS(const UpArgs &a)                                                                       \
    {                                                                                                    \
        const orc_elev_params *p = a.p;                                                                  \
        const int W = p->W;                                                                              \
        NS::tileWSDF = vec4((float) W, p->pixel_size, (float) p->grid, p->flip ? 1.0f : 0.0f);           \
        ref_sampler par, res, noi;                                                                       \
        par.w = par.h = W; par.channels = 3; par.filter = a.parent_filter; par.wrap = REF_CLAMP;         \
        par.subtexel_bits = a.subtexel_bits;                                                             \
        par.data = a.parent;                                                                             \
        if (a.parent != NULL) {                                                                          \
            float dx = (float) p->dx, dy = (float) p->dy;                                                \
            NS::coarseLevelOSL = vec4(dx / W, dy / W, 1.0f / W, 0.0f);                                   \
        } else {                                                                                         \
            NS::coarseLevelOSL = vec4(-1.0f, -1.0f, -1.0f, -1.0f);                                       \
        }                                                                                                \
        NS::coarseLevelSampler = par;                                                                    \
        /* the producer copies the W x W window of the residual tile and uploads it (:324-339) */        \
        float *win = new float[(size_t) W * W];                                                          \
        std::memset(win, 0, sizeof(float) * W * W);                                                      \
        if (p->has_resid && a.resid != NULL) {                                                           \
            for (int y = 0; y < W; ++y)                                                                  \
                for (int x = 0; x < W; ++x)                                                              \
                    win[x + y * W] = a.resid[(x + p->rx) + (y + p->ry) * p->resid_stride];               \
            NS::residualOSH = vec4(0.25f / W, 0.25f / W, 2.0f / W, 1.0f);                                \
        } else {                                                                                         \
            NS::residualOSH = vec4(0.0f, 0.0f, 1.0f, 0.0f);                                              \
        }                                                                                                \
        res.w = res.h = W; res.channels = 1; res.filter = REF_NEAREST; res.data = win;                   \
        NS::residualSampler = res;                                                                       \
        noi.w = noi.h = W; noi.layers = 6; noi.channels = 1; noi.filter = REF_NEAREST;                   \
        noi.wrap = REF_REPEAT; noi.data = a.noise;                                                       \
        NS::noiseSampler = noi;                                                                          \
        NS::noiseUVLH = vec4((float) p->noiseR, (float) ((p->noiseR + 1) % 4), (float) p->noiseL, p->rs); \
        for (int y = 0; y < W; ++y) {                                                                    \
            for (int x = 0; x < W; ++x) {                                                                \
                NS::st = vec2((float) x + 0.5f, (float) y + 0.5f);                                       \
                NS::shader_main();                                                                       \
                float *o = a.out + (size_t) (x + y * W) * 3;                                             \
                o[0] = NS::data.x; o[1] = NS::data.y; o[2] = NS::data.z;                                 \
            }                                                                                            \
        }                                                                                                \
        delete[] win;                                                                                    \
    }
UPSAMPLE_RUNNER(up_a)
UPSAMPLE_RUNNER(up_b)
UPSAMPLE_RUNNER(up_c)
UPSAMPLE_RUNNER(up_d)
UPSAMPLE_RUNNER(up_d_noclamp)

struct NrmArgs {
    const orc_norm_params *p;
    const float *elev;       /* elev_W^2 * 3 */
    const float *parent;     /* W*W*4 as the sampler returns it, or NULL */
    int subtexel_bits;
    float *out;              /* W*W*4 */
};

/* GLSL column c of a row-major maths matrix */
mat4 cols4(const float *m)
{
    mat4 r;
    for (int c = 0; c < 4; ++c) r[c] = vec4(m[0 * 4 + c], m[1 * 4 + c], m[2 * 4 + c], m[3 * 4 + c]);
    return r;
}
mat3 cols3(const float *m)
{
    mat3 r;
    for (int c = 0; c < 3; ++c) r[c] = vec3(m[0 * 3 + c], m[1 * 3 + c], m[2 * 3 + c]);
    return r;
}

/* NormalProducer.cpp:195-283 (the fp64 patch geometry arrives narrowed to fp32 in orc_norm_params,
 * exactly what setMatrix(... .cast<float>()) uploads) */
#define NORMAL_COMMON(NS)                                                                                \
        const orc_norm_params *p = a.p;                                                                  \
        const int W = p->W, EW = p->elev_W;                                                              \
        NS::tileSDF = vec3((float) W, (float) p->grid, (float) p->format);                               \
        ref_sampler el;                                                                                  \
        el.w = el.h = EW; el.channels = 3; el.filter = p->elev_filter; el.wrap = REF_CLAMP;              \
        el.subtexel_bits = a.subtexel_bits; el.data = a.elev;                                            \
        NS::elevationSampler = el;                                                                       \
        float bd = (float) p->elev_border;                                                               \
        NS::elevationOSL = vec4((bd + 0.25f) / EW, (bd + 0.25f) / EW, 1.0f / EW, 0.0f);                  \
        NS::deform = vec4(p->deform[0], p->deform[1], p->deform[2], p->deform[3]);
#define NORMAL_SPHERE_UNIFORMS(NS)                                                                       \
        NS::patchCorners = cols4(p->corners);                                                            \
        NS::patchVerticals = cols4(p->verticals);                                                        \
        NS::patchCornerNorms = vec4(p->norms[0], p->norms[1], p->norms[2], p->norms[3]);                 \
        NS::worldToTangentFrame = cols3(p->w2t);
#define NORMAL_LOOP(NS)                                                                                  \
        for (int y = 0; y < W; ++y) {                                                                    \
            for (int x = 0; x < W; ++x) {                                                                \
                NS::st = vec2((float) x + 0.5f, (float) y + 0.5f);                                       \
                NS::shader_main();                                                                       \
                float *o = a.out + (size_t) (x + y * W) * 4;                                             \
                o[0] = NS::data.x; o[1] = NS::data.y; o[2] = NS::data.z; o[3] = NS::data.w;              \
            }                                                                                            \
        }

void run_nrm_flat(const NrmArgs &a)
{
    NORMAL_COMMON(nrm_flat)
    NORMAL_LOOP(nrm_flat)
}
void run_nrm_sphere(const NrmArgs &a)
{
    NORMAL_COMMON(nrm_sphere)
    NORMAL_SPHERE_UNIFORMS(nrm_sphere)
    NORMAL_LOOP(nrm_sphere)
}
void run_nrm_demo(const NrmArgs &a)
{
    NORMAL_COMMON(nrm_demo)
    NORMAL_SPHERE_UNIFORMS(nrm_demo)
    nrm_demo::parentToTangentFrame = cols3(p->p2t);
    ref_sampler pn;
    pn.w = pn.h = W; pn.channels = 4; pn.filter = p->parent_filter; pn.wrap = REF_CLAMP;
    pn.subtexel_bits = a.subtexel_bits; pn.data = a.parent;
    nrm_demo::normalSampler = pn;
    if (p->has_parent && a.parent != NULL) {
        float dx = (float) p->ptx * ((float) W / 2.0f);
        float dy = (float) p->pty * ((float) W / 2.0f);
        nrm_demo::normalOSL = vec4((dx + 0.25f) / W, (dy + 0.25f) / W, 1.0f / W, 0.0f);
    } else {
        nrm_demo::normalOSL = vec4(-1.0f, -1.0f, -1.0f, -1.0f);
    }
    NORMAL_LOOP(nrm_demo)
}

struct OrthoArgs {
    const orc_ortho_params *p;
    const uint8_t *parent;   /* W*W*4 or NULL */
    int parent_filter;
    const uint8_t *residual; /* W*W*channels or NULL */
    int channels;
    const uint8_t *noise;    /* 6*W*W*4 */
    float *out;              /* W*W*4: `data` before the colour-buffer conversion */
};

/* OrthoProducer.cpp:298-372 */
#define ORTHO_RUNNER(NS)                                                                                 \
    void run_##NS(const OrthoArgs &a)                                                                    \
    {                                                                                                    \
        const orc_ortho_params *p = a.p;                                                                 \
        const int W = p->tileWidth;                                                                      \
        NS::tileWidth = (float) W;                                                                       \
        ref_sampler par, res, noi;                                                                       \
        par.w = par.h = W; par.channels = 4; par.filter = a.parent_filter; par.bytes = a.parent;         \
        NS::coarseLevelSampler = par;                                                                    \
        if (p->level > 0 && a.parent != NULL) {                                                          \
            float dx = (float) p->dx, dy = (float) p->dy;                                                \
            NS::coarseLevelOSL = vec4((dx + 0.5f) / W, (dy + 0.5f) / W, 1.0f / W, 0.0f);                 \
        } else {                                                                                         \
            NS::coarseLevelOSL = vec4(-1.0f, -1.0f, -1.0f, -1.0f);                                       \
        }                                                                                                \
        res.w = res.h = W; res.channels = a.channels; res.filter = REF_NEAREST; res.bytes = a.residual;  \
        NS::residualSampler = res;                                                                       \
        if (p->hasResidual && a.residual != NULL) {                                                      \
            NS::residualOSH = vec4(0.5f / W, 0.5f / W, 1.0f / W, p->residualScale);                      \
        } else {                                                                                         \
            NS::residualOSH = vec4(-1.0f, -1.0f, -1.0f, -1.0f);                                          \
        }                                                                                                \
        noi.w = noi.h = W; noi.layers = 6; noi.channels = 4; noi.filter = REF_NEAREST;                   \
        noi.wrap = REF_REPEAT; noi.bytes = a.noise;                                                      \
        NS::noiseSampler = noi;                                                                          \
        NS::noiseUVLH = ivec4(p->noiseR, (p->noiseR + 1) % 4, p->noiseL, p->hsv ? 1 : 0);                \
        NS::noiseColor = vec4(p->noiseColor[0], p->noiseColor[1], p->noiseColor[2], p->noiseColor[3]);   \
        NS::rootNoiseColor = vec4(p->rootNoiseColor[0], p->rootNoiseColor[1], p->rootNoiseColor[2],      \
                                  p->rootNoiseColor[3]);                                                 \
        for (int y = 0; y < W; ++y) {                                                                    \
            for (int x = 0; x < W; ++x) {                                                                \
                NS::st = vec2((float) x + 0.5f, (float) y + 0.5f);                                       \
                NS::shader_main();                                                                       \
                float *o = a.out + (size_t) (x + y * W) * 4;                                             \
                o[0] = NS::data.x; o[1] = NS::data.y; o[2] = NS::data.z; o[3] = NS::data.w;              \
            }                                                                                            \
        }                                                                                                \
    }
ORTHO_RUNNER(ortho_demo)
ORTHO_RUNNER(ortho_ex)

} /* namespace */

extern "C" {

/* variant: 0 = A (terrain1), 1 = B (terrain2), 2 = C (terrain4), 3 = D (demo), 4 = D with NO_CLAMP.
 * parent_filter: min/mag filter of the elevation storage (0 NEAREST, 1 LINEAR); subtexel_bits: see the shim. */
int ref_glsl_upsample_tile(int variant, const orc_elev_params *p, const float *parent, int parent_filter,
                           const float *resid, const float *noise, int subtexel_bits, float *result)
{
    UpArgs a = { p, parent, parent_filter, resid, noise, subtexel_bits, result };
    switch (variant) {
    case 0: run_up_a(a); return 0;
    case 1: run_up_b(a); return 0;
    case 2: run_up_c(a); return 0;
    case 3: run_up_d(a); return 0;
    case 4: run_up_d_noclamp(a); return 0;
    }
    return -1;
}

/* variant: 0 = terrain1 (flat, RG8), 1 = terrain2 (sphere, RG8), 2 = demo */
int ref_glsl_normal_tile(int variant, const orc_norm_params *p, const float *elev, const float *parent,
                         int subtexel_bits, float *result)
{
    NrmArgs a = { p, elev, parent, subtexel_bits, result };
    switch (variant) {
    case 0: run_nrm_flat(a); return 0;
    case 1: run_nrm_sphere(a); return 0;
    case 2: run_nrm_demo(a); return 0;
    }
    return -1;
}

/* variant: 0 = demo, 1 = the examples' copy */
int ref_glsl_ortho_tile(int variant, const orc_ortho_params *p, const uint8_t *parent, int parent_filter,
                        const uint8_t *residual, int channels, const uint8_t *noise, float *result)
{
    OrthoArgs a = { p, parent, parent_filter, residual, channels, noise, result };
    switch (variant) {
    case 0: run_ortho_demo(a); return 0;
    case 1: run_ortho_ex(a); return 0;
    }
    return -1;
}

/* The quadtree of a scene, produced by the reference's own shader text: per tile the uniforms of ElevationProducer /
 * NormalProducer::doCreateTile (the restatement's: orc_elev_uniforms / orc_normal_uniforms, liborc.so -- the reference's
 * host code needs Ork), then upsampleShader.glsl (demo variant: slope noise, clamp or NO_CLAMP) and normalShader.glsl (demo)
 * over every fragment, then the RG8 packing.  Tiles of a level in parallel (OpenMP; the shader globals are thread_local).
 * The same signature and results layout as orc_produce_quadtree: bench.py's `--impl reference` arm and cpu_baseline run
 * this when oracle/_ref exists (kind "reference"); scenes without residuals. */
long ref_glsl_produce_quadtree(const orc_scene *s, int maxLevel, int nthreads, double *checksum, float *zmin, float *zmax)
{
    const int W = s->W, NW = W - 4;
    const size_t esz = (size_t) W * W * 3;
    if (s->resid != NULL || s->noise_mode != 1) return -1;        /* the demo variant: slope-modulated noise */
    std::vector<float> noise((size_t) 6 * W * W);
    orc_dem_noise_r16f(W, noise.data());
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    std::vector<float> prev, cur;
    long produced = 0;
    double sum = 0.0;
    float lo = INFINITY, hi = -INFINITY;
    for (int level = 0; level <= maxLevel; ++level) {
        const long n = 1L << level, count = n * n;
        cur.assign(esz * (size_t) count, 0.0f);
        double lsum = 0.0;
        float llo = INFINITY, lhi = -INFINITY;
#pragma omp parallel
        {
            std::vector<float> data((size_t) NW * NW * 4);
            std::vector<uint8_t> norm((size_t) NW * NW * 2);
#pragma omp for schedule(dynamic, 4) reduction(+ : lsum) reduction(min : llo) reduction(max : lhi)
            for (long t = 0; t < count; ++t) {
                const int tx = (int) (t % n), ty = (int) (t / n);
                const float *parent = level > 0 ? prev.data() + esz * ((size_t) (tx / 2) + (size_t) (ty / 2) * (n / 2)) : NULL;
                float *e = cur.data() + esz * (size_t) t;
                orc_elev_params ep;
                orc_elev_uniforms(W, s->gridMeshSize, s->rootQuadSize, s->flip, s->noiseAmp, s->nAmp, s->face, level, tx, ty, 0, 0,
                                  s->noise_mode, s->no_clamp, &ep);
                ref_glsl_upsample_tile(s->no_clamp ? 4 : 3, &ep, parent, s->elev_filter, NULL, noise.data(), 8, e);
                orc_norm_params np;
                orc_normal_uniforms(NW, s->gridMeshSize, 2, 0, W, 2, s->elev_filter, ORC_FILTER_LINEAR, (double) s->rootQuadSize,
                                    s->sphere, level, tx, ty, &np);
                ref_glsl_normal_tile(2, &np, e, NULL, 8, data.data());
                orc_pack_unorm8(NW, 2, data.data(), norm.data());
                float a, b;
                orc_tile_minmax(W, e, &a, &b);
                lsum += (double) a + (double) b;
                llo = fminf(llo, a);
                lhi = fmaxf(lhi, b);
            }
        }
        produced += count;
        sum += lsum;
        lo = fminf(lo, llo);
        hi = fmaxf(hi, lhi);
        prev.swap(cur);
    }
    if (checksum) *checksum = sum;
    if (zmin) *zmin = lo;
    if (zmax) *zmax = hi;
    return produced;
}

} /* extern "C" */
