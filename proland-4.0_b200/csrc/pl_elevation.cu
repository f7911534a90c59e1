/*
 * pl_elevation.cu -- the batched ElevationProducer pass for sm_100a.
 *
 * One CTA produces one tile: what the reference draws as one full-screen quad
 * with upsampleShader (src/demo/shaders/elevation/upsampleShader.glsl:140-203,
 * variants A/B/C in src/terrain/examples/terrain{1,2,4}/upsampleShader.glsl)
 * after ElevationProducer::doCreateTile (ElevationProducer.cpp:280-405) has set
 * the uniforms.  A batch is a grid of such CTAs; siblings are neighbours in the
 * grid so the parent tile they share is read from HBM once and from L2 after.
 *
 *   - the parent's zf window ((tileSize/2+6)^2 texels at (dx,dy)) is staged in
 *     shared memory by ONE TMA tensor copy (cp.async.bulk.tensor.3d) signalled
 *     on an mbarrier; texels past the tile edge are zero-filled and only ever
 *     meet weight 0
 *   - the sparse parent zm samples of the coarse height zc sit on a lattice of
 *     pitch `grid`; they are gathered once per tile into shared memory with the
 *     clamp-to-edge the GL sampler applies
 *   - a thread owns a 2x2 output quad: the four parities share the same 4x4
 *     parent neighbourhood (16 LDS for 4 texels x 3 channels)
 *   - noise comes from the pre-rotated fp16 planes (pl_ctx.cu): one half2 load
 *     per quad row, coalesced whatever the rotation
 *   - outputs go straight from registers to the three planes as float2 stores;
 *     pad columns are written too, so every 32-byte sector is written whole
 *
 * Arithmetic: fp32 in the canonical order of oracle/orc_fp.h -- this file is
 * compiled with --fmad=false and spells every fused multiply-add as fmaf(), so
 * the result is bit-identical to the CPU oracle.  Zero-weight taps of the GLSL
 * constant matrices are dropped (value-identical for finite inputs).
 */
#include "pl_elevation_tile.cuh"

namespace {

using namespace plelev;

constexpr int kThreads = 192;

/* ------------------------------------------------------------------------
 * Generic kernel: any odd tile width, any grid divisor (runtime geometry).
 * Used for everything but the shipped geometry (tile 101, grid 4), which has
 * the specialised kernel below.
 * ------------------------------------------------------------------------ */
template <int RESID>
__global__ void __launch_bounds__(kThreads) elevation_kernel_generic(const __grid_constant__ CUtensorMap tm, const ElevArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *win = reinterpret_cast<float *>(smem_raw);
    float *lat = win + a.box_h * a.box_w;
    uchar4 *lut = reinterpret_cast<uchar4 *>(lat + a.nk * a.nk);
    __shared__ uint64_t bar;
    __shared__ float red_lo[kThreads / 32], red_hi[kThreads / 32];

    const int tid = threadIdx.x;
    const pl_elev_req rq = a.reqs[blockIdx.x];
    const bool has_parent = rq.parent_slot >= 0;
    const int W = a.W, g = a.grid;

    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (has_parent) {
            mbar_expect_tx(&bar, (uint32_t) (a.box_w * a.box_h * sizeof(float)));
            tma_load_3d(win, &tm, &bar, rq.dx, rq.dy, rq.parent_slot * 3);
        }
    }

    /* per-column lattice indices: floor(ij/2g + 0.5) and floor(ij/2g), +1 so that -1 -> 0 */
    for (int x = tid; x < W; x += kThreads) {
        const int ij = x - 2;
        const int kr = floordiv(ij + g, 2 * g) + 1;
        const int kf = floordiv(ij, 2 * g) + 1;
        int m = ij % (2 * g);
        if (m < 0) m += 2 * g;
        lut[x] = make_uchar4((unsigned char) kr, (unsigned char) kf, (unsigned char) (m == g), 0);
    }
    if (has_parent) {
        const float *pzm = a.elev + (size_t) rq.parent_slot * 3 * a.plane + 2 * (size_t) a.plane;
        for (int k = tid; k < a.nk * a.nk; k += kThreads) {
            const int kx = k % a.nk - 1, ky = k / a.nk - 1;
            const int px = min(max(2 + g * kx + rq.dx, 0), W - 1);   /* CLAMP_TO_EDGE */
            const int py = min(max(2 + g * ky + rq.dy, 0), W - 1);
            lat[k] = __ldg(pzm + (size_t) py * a.pitch + px);
        }
    } else {
        for (int k = tid; k < a.box_w * a.box_h; k += kThreads) win[k] = 0.0f;
    }
    __syncthreads();
    if (has_parent) mbar_wait(&bar, 0);

    const float rs = rq.rs;
    const float pixel = rq.pixel_size;
    const float nvz = 2.0f * pixel;
    /* tile-uniform noise path; rs == 0 adds exactly 0 in every variant */
    const int nz = rs == 0.0f ? NZ_NONE : (a.noise_mode == PL_NOISE_PLAIN ? NZ_PLAIN : (rs < 0.0f ? NZ_NEG : NZ_SLOPE));
    const int flip = a.flip;
    const float ars = fabsf(rs);

    float *out = a.elev + (size_t) rq.out_slot * 3 * a.plane;
    const __half *nplane = a.noise + (size_t) (rq.noise_r * 6 + rq.noise_l) * a.noise_plane;
    const int QW = a.pitch >> 1, QH = (W + 1) >> 1;
    const int bw = a.box_w;

    float lo = INFINITY, hi = -INFINITY;

    for (int it = tid; it < QW * QH; it += kThreads) {
        const int j = it / QW, i = it - j * QW;
        const int x0 = 2 * i, y0 = 2 * j;
        const bool vy = y0 + 1 < W;
        float *o0 = out + (size_t) y0 * a.pitch + x0;
        float *o1 = o0 + a.pitch;
        if (x0 >= W) {   /* pad quad: keeps the sectors of the row whole */
            const float2 z2 = make_float2(0.0f, 0.0f);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                *reinterpret_cast<float2 *>(o0 + (size_t) c * a.plane) = z2;
                if (vy) *reinterpret_cast<float2 *>(o1 + (size_t) c * a.plane) = z2;
            }
            continue;
        }
        const bool vx = x0 + 1 < W;

        /* residual (residualOSH.w * texel) */
        float r00 = 0.0f, r10 = 0.0f, r01 = 0.0f, r11 = 0.0f;   /* r{x}{y} */
        if (RESID != 0 && rq.resid_slot >= 0) {
            const size_t off = (size_t) rq.resid_slot * a.resid_slot_elems + (size_t) (y0 + rq.ry) * a.resid_pitch + (x0 + rq.rx);
            if (RESID == 1) {
                const float *rp = static_cast<const float *>(a.resid) + off;
                const float2 t0 = __ldg(reinterpret_cast<const float2 *>(rp));
                r00 = t0.x; r10 = t0.y;
                if (vy) { const float2 t1 = __ldg(reinterpret_cast<const float2 *>(rp + a.resid_pitch)); r01 = t1.x; r11 = t1.y; }
            } else {
                const short *rp = static_cast<const short *>(a.resid) + off;
                const short2 t0 = __ldg(reinterpret_cast<const short2 *>(rp));
                r00 = (float) t0.x * a.resid_scale; r10 = (float) t0.y * a.resid_scale;
                if (vy) {
                    const short2 t1 = __ldg(reinterpret_cast<const short2 *>(rp + a.resid_pitch));
                    r01 = (float) t1.x * a.resid_scale; r11 = (float) t1.y * a.resid_scale;
                }
            }
        }

        /* cz[c][r] = parent.zf[bx + r, by + c]; z<c><r> */
        const float *w = win + j * bw + i;
        const float z00 = w[0], z01 = w[1], z02 = w[2], z03 = w[3];
        const float z10 = w[bw], z11 = w[bw + 1], z12 = w[bw + 2], z13 = w[bw + 3];
        const float z20 = w[2 * bw], z21 = w[2 * bw + 1], z22 = w[2 * bw + 2], z23 = w[2 * bw + 3];
        const float z30 = w[3 * bw], z31 = w[3 * bw + 1], z32 = w[3 * bw + 2], z33 = w[3 * bw + 3];
        (void) z00; (void) z03; (void) z30; (void) z33;

        /* noise texels (already rotated) */
        float n00 = 0.0f, n10 = 0.0f, n01 = 0.0f, n11 = 0.0f;
        if (nz != NZ_NONE) {
            /* columns are permuted inside groups of four (pl_noise_col): x0 even -> x0 + 1 sits two further */
            const __half *np = nplane + (size_t) y0 * a.noise_pitch + pl_noise_col(x0);
            n00 = __half2float(__ldg(np)); n10 = __half2float(__ldg(np + 2));
            if (vy) { n01 = __half2float(__ldg(np + a.noise_pitch)); n11 = __half2float(__ldg(np + a.noise_pitch + 2)); }
        }

        /* zf before the upsample term */
        float f00 = r00, f10 = r10, f01 = r01, f11 = r11;
        if (nz == NZ_PLAIN) {
            f00 = fmaf(ars, n00, r00); f10 = fmaf(ars, n10, r10); f01 = fmaf(ars, n01, r01); f11 = fmaf(ars, n11, r11);
        } else if (nz == NZ_NEG) {
            f00 = fmaf(-rs, n00, r00); f10 = fmaf(-rs, n10, r10); f01 = fmaf(-rs, n01, r01); f11 = fmaf(-rs, n11, r11);
        } else if (nz == NZ_SLOPE) {
            /* parity 0 (x even, y even): slopex/slopey/curvature matrices [0] */
            {
                const float sx = z10 - z12;
                const float sy = z01 - z21;
                const float cv = ((-z01) + (fmaf(z11, 4.0f, -z10) - z12)) + (-z21);
                f00 = fmaf(noise_amp(sx, sy, cv, nvz, pixel) * rs, n00, r00);
            }
            /* parity 1 (x odd, y even): [1] */
            {
                const float sx = chain4(z10, z11, z12, z13, 0.5f, 0.5f, -0.5f, -0.5f);
                const float sy = fmaf(z02, 0.5f, z01 * 0.5f) + fmaf(z22, -0.5f, z21 * -0.5f);
                const float cv = (fmaf(z02, -0.5f, z01 * -0.5f) + chain4(z10, z11, z12, z13, -0.5f, 1.5f, 1.5f, -0.5f))
                                 + fmaf(z22, -0.5f, z21 * -0.5f);
                f10 = fmaf(noise_amp(sx, sy, cv, nvz, pixel) * rs, n10, r10);
            }
            /* parity 2 (x even, y odd): [2] */
            {
                const float sx = fmaf(z12, -0.5f, z10 * 0.5f) + fmaf(z22, -0.5f, z20 * 0.5f);
                const float sy = ((z01 * 0.5f + z11 * 0.5f) + z21 * -0.5f) + z31 * -0.5f;
                const float cv = ((z01 * -0.5f + fmaf(z12, -0.5f, fmaf(z11, 1.5f, z10 * -0.5f)))
                                  + fmaf(z22, -0.5f, fmaf(z21, 1.5f, z20 * -0.5f))) + z31 * -0.5f;
                f01 = fmaf(noise_amp(sx, sy, cv, nvz, pixel) * rs, n01, r01);
            }
            /* parity 3 (x odd, y odd): [3] */
            {
                const float sx = chain4(z10, z11, z12, z13, 0.25f, 0.25f, -0.25f, -0.25f)
                                 + chain4(z20, z21, z22, z23, 0.25f, 0.25f, -0.25f, -0.25f);
                const float sy = ((fmaf(z02, 0.25f, z01 * 0.25f) + fmaf(z12, 0.25f, z11 * 0.25f))
                                  + fmaf(z22, -0.25f, z21 * -0.25f)) + fmaf(z32, -0.25f, z31 * -0.25f);
                const float cv = ((fmaf(z02, -0.25f, z01 * -0.25f) + chain4(z10, z11, z12, z13, -0.25f, 0.5f, 0.5f, -0.25f))
                                  + chain4(z20, z21, z22, z23, -0.25f, 0.5f, 0.5f, -0.25f)) + fmaf(z32, -0.25f, z31 * -0.25f);
                f11 = fmaf(noise_amp(sx, sy, cv, nvz, pixel) * rs, n11, r11);
            }
        }

        float c00 = f00, c10 = f10, c01 = f01, c11 = f11;   /* zc (level 0: zc = zf) */
        if (has_parent) {
            /* upsampleMatrix[0..3] */
            const float W1 = -1.0f / 16.0f, W9 = 9.0f / 16.0f;
            const float V1 = 1.0f / 256.0f, V9 = -9.0f / 256.0f, V81 = 81.0f / 256.0f;
            f00 = f00 + z11;
            f10 = f10 + chain4(z10, z11, z12, z13, W1, W9, W9, W1);
            f01 = f01 + (((z01 * W1 + z11 * W9) + z21 * W9) + z31 * W1);
            f11 = f11 + (((chain4(z00, z01, z02, z03, V1, V9, V9, V1) + chain4(z10, z11, z12, z13, V9, V81, V81, V9))
                          + chain4(z20, z21, z22, z23, V9, V81, V81, V9)) + chain4(z30, z31, z32, z33, V1, V9, V9, V1));

            /* coarse height: zc1 = zm[round_x, floor_y], zc3 = zm[floor_x, round_y] */
            const uchar4 lx0 = lut[x0], lx1 = lut[vx ? x0 + 1 : x0];
            const uchar4 ly0 = lut[y0], ly1 = lut[vy ? y0 + 1 : y0];
            const int nk = a.nk;
#define ZC(LX, LY, OUT)                                                                             \
            {                                                                                       \
                const float zc1 = lat[LX.x + LY.y * nk], zc3 = lat[LX.y + LY.x * nk];               \
                if (flip && LX.z && LY.z) {                                                         \
                    const float zc0 = lat[LX.y + LY.y * nk], zc2 = lat[LX.x + LY.x * nk];           \
                    OUT = (zc3 + zc1 >= zc0 + zc2 ? zc1 + zc3 : zc0 + zc2) * 0.5f;                  \
                } else {                                                                            \
                    OUT = (zc1 + zc3) * 0.5f;                                                       \
                }                                                                                   \
            }
            ZC(lx0, ly0, c00) ZC(lx1, ly0, c10) ZC(lx0, ly1, c01) ZC(lx1, ly1, c11)
#undef ZC
        }

        float m00 = f00, m10 = f10, m01 = f01, m11 = f11;
        if (!a.no_clamp) { m00 = fmaxf(f00, 0.0f); m10 = fmaxf(f10, 0.0f); m01 = fmaxf(f01, 0.0f); m11 = fmaxf(f11, 0.0f); }

        if (a.want_stats) {   /* TileSamplerZ.cpp:60-64: texels [2, W-3]^2 of zm */
            const bool ix0 = x0 >= 2 && x0 <= W - 3, ix1 = x0 + 1 >= 2 && x0 + 1 <= W - 3;
            const bool iy0 = y0 >= 2 && y0 <= W - 3, iy1 = y0 + 1 >= 2 && y0 + 1 <= W - 3;
            if (ix0 && iy0) { lo = fminf(lo, m00); hi = fmaxf(hi, m00); }
            if (ix1 && iy0) { lo = fminf(lo, m10); hi = fmaxf(hi, m10); }
            if (ix0 && iy1) { lo = fminf(lo, m01); hi = fmaxf(hi, m01); }
            if (ix1 && iy1) { lo = fminf(lo, m11); hi = fmaxf(hi, m11); }
        }

        if (!vx) { f10 = c10 = m10 = 0.0f; f11 = c11 = m11 = 0.0f; }   /* pad column */
        *reinterpret_cast<float2 *>(o0) = make_float2(f00, f10);
        *reinterpret_cast<float2 *>(o0 + a.plane) = make_float2(c00, c10);
        *reinterpret_cast<float2 *>(o0 + 2 * (size_t) a.plane) = make_float2(m00, m10);
        if (vy) {
            *reinterpret_cast<float2 *>(o1) = make_float2(f01, f11);
            *reinterpret_cast<float2 *>(o1 + a.plane) = make_float2(c01, c11);
            *reinterpret_cast<float2 *>(o1 + 2 * (size_t) a.plane) = make_float2(m01, m11);
        }
    }

    if (a.want_stats) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, s));
            hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, s));
        }
        if ((tid & 31) == 0) { red_lo[tid >> 5] = lo; red_hi[tid >> 5] = hi; }
        __syncthreads();
        if (tid == 0) {
#pragma unroll
            for (int k = 1; k < kThreads / 32; ++k) { lo = fminf(lo, red_lo[k]); hi = fmaxf(hi, red_hi[k]); }
            a.stats[rq.out_slot] = make_float2(lo, hi);
        }
    }
}

/* ------------------------------------------------------------------------
 * Specialised kernel: compile-time geometry, one CTA per tile; the per-tile
 * device code lives in pl_elevation_tile.cuh (shared with pl_pair.cu).
 * ------------------------------------------------------------------------ */
template <int TW, int TG, int RESID>
__global__ void __launch_bounds__(kThreads) elevation_kernel_fast(const __grid_constant__ CUtensorMap tm, const ElevArgs a)
{
    __shared__ __align__(128) float smem[ElevSmem<TW, TG>::FLOATS];
    __shared__ uint64_t bar;
    const int tid = threadIdx.x;
    const pl_elev_req rq = a.reqs[blockIdx.x];
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    elevation_tile<TW, TG, RESID, kThreads, false>(&tm, a, rq, smem, &bar, 0, nullptr, tid);
}

}  // namespace


int pl_elev_fill_args(pl_ctx *ctx, const pl_elev_scene *sc, pl_pool *elev, pl_pool *resid, const pl_elev_req *dev_reqs,
                      plelev::ElevArgs &a)
{
    a.elev = reinterpret_cast<float *>(elev->base);
    a.resid = resid ? resid->base : nullptr;
    const pl_ctx::NoiseTable *nt = ctx->noise_for(sc->tile_w);
    if (!nt) return pl_set_error(PL_ERR_ARG, "pl_noise_init(ctx, %d) has not been called", sc->tile_w);
    a.noise = nt->rot;
    a.reqs = dev_reqs;
    a.stats = elev->stats;
    a.ready = nullptr;
    a.epoch = 0;
    a.W = elev->tile_w;
    a.pitch = elev->pitch;
    a.plane = (int) elev->plane_elems;
    a.grid = sc->grid;
    a.flip = sc->flip ? 1 : 0;
    a.noise_mode = sc->noise_mode;
    a.no_clamp = sc->no_clamp;
    a.want_stats = sc->want_stats ? 1 : 0;
    a.box_w = elev->box_w;
    a.box_h = elev->box_h;
    /* lattice indices k in [-1, floor((W-3+g)/2g)] */
    a.nk = (a.W - 3 + a.grid) / (2 * a.grid) + 2;
    if (a.nk > 255) return pl_set_error(PL_ERR_ARG, "grid %d too fine for tile_w %d", a.grid, a.W);
    a.noise_pitch = nt->pitch;
    a.noise_plane = nt->w * nt->pitch;
    a.resid_pitch = resid ? resid->pitch : 0;
    a.resid_slot_elems = resid ? (long long) (resid->slot_bytes / (resid->kind == PL_POOL_RESID_F32 ? 4 : 2)) : 0;
    a.resid_scale = sc->resid_scale;
    return PL_OK;
}

int pl_launch_elevation(pl_ctx *ctx, const pl_elev_scene *sc, pl_pool *elev, pl_pool *resid, int n,
                        const pl_elev_req *dev_reqs)
{
    PL_CUDA(cudaSetDevice(ctx->device));
    ElevArgs a;
    int rc = pl_elev_fill_args(ctx, sc, elev, resid, dev_reqs, a);
    if (rc) return rc;

    const size_t smem = (size_t) a.box_w * a.box_h * 4 + (size_t) a.nk * a.nk * 4 + (size_t) a.W * 4;
    const int rk = !resid ? 0 : (resid->kind == PL_POOL_RESID_F32 ? 1 : 2);
    if (a.W == 101 && a.grid == 4 && !ctx->force_generic) {
        /* the geometry of every shipped archive: compile-time specialisation */
        auto kern = rk == 0 ? elevation_kernel_fast<101, 4, 0>
                            : (rk == 1 ? elevation_kernel_fast<101, 4, 1> : elevation_kernel_fast<101, 4, 2>);
        pl_timing_begin(ctx, PL_K_ELEVATION, n);
        kern<<<n, kThreads, 0, ctx->stream>>>(elev->tm_parent, a);
        pl_timing_end(ctx);
    } else {
        auto kern = rk == 0 ? elevation_kernel_generic<0>
                            : (rk == 1 ? elevation_kernel_generic<1> : elevation_kernel_generic<2>);
        if (smem > 40 * 1024) PL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        pl_timing_begin(ctx, PL_K_ELEVATION, n);
        kern<<<n, kThreads, smem, ctx->stream>>>(elev->tm_parent, a);
        pl_timing_end(ctx);
    }
    PL_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return PL_OK;
}
