/*
 * orc_driver.c -- ORACLE (test infrastructure, never shipped): produces
 * elevation + normal tile pairs over a quadtree the way the reference's
 * producers chain them: parent elevation -> elevation ->
 * normal.  Used by the parity tests and as bench.py's cpu_baseline /
 * --impl reference arm (kind "port": the reference's GL path cannot be built
 * or run here).
 *
 * Restates the wiring of
 *   terrain/sources/proland/dem/ElevationProducer.cpp:246-271,280-405
 *   terrain/sources/proland/dem/NormalProducer.cpp:139-156,164-289
 */
#include "orc.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

void orc_produce_pair(const orc_scene *s, const float *noise, int level, int tx, int ty,
                      const float *parent, const float *resid_tile,
                      float *elev_out, uint8_t *norm_out)
{
    const int W = s->W;
    const int NW = W - 4;
    orc_elev_params ep;
    int resid_W = s->resid ? s->resid->tileSize + 5 : 0;
    orc_elev_uniforms(W, s->gridMeshSize, s->rootQuadSize, s->flip, s->noiseAmp, s->nAmp,
                      s->face, level, tx, ty, resid_tile != NULL, resid_W,
                      s->noise_mode, s->no_clamp, &ep);
    orc_upsample_tile(&ep, level > 0 ? parent : NULL, resid_tile, noise, elev_out);

    if (norm_out) {
        orc_norm_params np;
        orc_normal_uniforms(NW, s->gridMeshSize, 2, 0, W, 2, s->elev_filter, ORC_FILTER_LINEAR,
                            (double) s->rootQuadSize, s->sphere, level, tx, ty, &np);
        float *data = (float *) malloc(sizeof(float) * NW * NW * 4);
        orc_normal_tile(&np, elev_out, NULL, data);
        orc_pack_unorm8(NW, 2, data, norm_out);
        free(data);
    }
}

long orc_produce_quadtree(const orc_scene *s, int maxLevel, int nthreads,
                          double *checksum, float *zmin, float *zmax)
{
    const int W = s->W;
    const int NW = W - 4;
    const size_t esz = (size_t) W * W * 3;
    const size_t nsz = (size_t) NW * NW * 2;
    float *noise = (float *) malloc(sizeof(float) * 6 * W * W);
    orc_dem_noise_r16f(W, noise);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void) nthreads;
#endif
    float *prev = NULL;
    long produced = 0;
    double sum = 0.0;
    float lo = INFINITY, hi = -INFINITY;
    const int resid_W = s->resid ? s->resid->tileSize + 5 : 0;
    const int mod = s->resid ? (resid_W - 5) / (W - 5) : 1;

    for (int level = 0; level <= maxLevel; ++level) {
        const long n = 1L << level;
        const long count = n * n;
        float *cur = (float *) malloc(sizeof(float) * esz * count);
        double lsum = 0.0;
        float llo = INFINITY, lhi = -INFINITY;
#pragma omp parallel
        {
            uint8_t *norm = (uint8_t *) malloc(nsz);
            float *rtile = s->resid ? (float *) malloc(sizeof(float) * resid_W * resid_W) : NULL;
#pragma omp for schedule(dynamic, 4) reduction(+ : lsum) reduction(min : llo) reduction(max : lhi)
            for (long t = 0; t < count; ++t) {
                const int tx = (int) (t % n), ty = (int) (t / n);
                const float *parent = level > 0
                    ? prev + esz * ((size_t) (tx / 2) + (size_t) (ty / 2) * (n / 2)) : NULL;
                const float *r = NULL;
                if (s->resid && orc_resid_has_tile(s->resid, level, tx / mod, ty / mod)) {
                    memset(rtile, 0, sizeof(float) * resid_W * resid_W);
                    orc_resid_create_tile(s->resid, level, tx / mod, ty / mod, rtile);
                    r = rtile;
                }
                float *e = cur + esz * (size_t) t;
                orc_produce_pair(s, noise, level, tx, ty, parent, r, e, norm);
                float a, b;
                orc_tile_minmax(W, e, &a, &b);
                lsum += (double) a + (double) b;
                llo = fminf(llo, a);
                lhi = fmaxf(lhi, b);
            }
            free(norm);
            free(rtile);
        }
        produced += count;
        sum += lsum;
        lo = fminf(lo, llo);
        hi = fmaxf(hi, lhi);
        free(prev);
        prev = cur;
    }
    free(prev);
    free(noise);
    if (checksum) *checksum = sum;
    if (zmin) *zmin = lo;
    if (zmax) *zmax = hi;
    return produced;
}
