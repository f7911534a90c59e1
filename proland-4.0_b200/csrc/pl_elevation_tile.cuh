/*
 * pl_elevation_tile.cuh -- device code of the ElevationProducer pass for ONE tile of the shipped
 * geometry (compile-time tile width / grid divisor), shared by the elevation kernel
 * (pl_elevation.cu) and the fused elevation+normal kernel (pl_pair.cu).
 *
 * Reference: upsampleShader.glsl:140-203 after ElevationProducer::doCreateTile
 * (ElevationProducer.cpp:280-405) has set the uniforms.  Arithmetic: canonical order of
 * oracle/orc_fp.h (compile with --fmad=false; every fused operation is spelled fmaf / fma2).
 */
#ifndef PL_ELEVATION_TILE_CUH
#define PL_ELEVATION_TILE_CUH

#include "pl_internal.h"
#include "pl_fpexact.cuh"
#include "pl_f2.cuh"

namespace plelev {


struct ElevArgs {
    float *elev;
    const void *resid;
    const __half *noise;
    const pl_elev_req *reqs;
    float2 *stats;
    int *ready;              /* pl_produce_levels: per-slot ready flags, NULL otherwise */
    int epoch;
    int W, pitch, plane;
    int grid, flip, noise_mode, no_clamp, want_stats;
    int box_w, box_h, nk;
    int noise_pitch, noise_plane;
    int resid_pitch;
    long long resid_slot_elems;
    float resid_scale;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *tm, uint64_t *bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t) tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

/* pl_produce_levels: tiles of several quadtree levels in ONE launch.  A tile whose request says so (pad_[0] != 0)
 * waits until its parent -- a CTA with a lower block index of the same launch -- has published its elevation planes:
 * the parent's threads fence their stores, meet, and one of them releases ready[slot] = epoch; the child's thread 0
 * acquires it, orders the generic-proxy view before its TMA copy (fence.proxy.async) and the CTA proceeds. */
__device__ __forceinline__ void levels_wait_parent(const ElevArgs &a, const pl_elev_req &rq, const int tid)
{
    if (a.ready == nullptr) return;
    if (tid == 0 && rq.pad_[0] != 0 && rq.parent_slot >= 0) {
        const int *flag = a.ready + rq.parent_slot;
        int v;
        do {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
            if (v != a.epoch) __nanosleep(64);
        } while (v != a.epoch);
        asm volatile("fence.proxy.async;" ::: "memory");
    }
    __syncthreads();
}
/* after the elevation planes of the tile have been stored: every thread calls this (contains a barrier) */
__device__ __forceinline__ void levels_publish(const ElevArgs &a, const pl_elev_req &rq, const int tid)
{
    if (a.ready == nullptr) return;
    __threadfence();
    __syncthreads();
    if (tid == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(a.ready + rq.out_slot), "r"(a.epoch) : "memory");
}

__device__ __forceinline__ int floordiv(int a, int b) { int q = a / b; return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q; }
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

/* dot(cz[c], w) with all four weights non-zero: canonical left-to-right chain */
__device__ __forceinline__ float chain4(float a0, float a1, float a2, float a3, float w0, float w1, float w2, float w3)
{
    return fmaf(a3, w3, fmaf(a2, w2, fmaf(a1, w1, a0 * w0)));
}

/* noise amplitude factor of variants B / D: upsampleShader.glsl:166-170 */
__device__ __forceinline__ float noise_amp(float nvx, float nvy, float curv_num, float nvz, float pixel)
{
    const float slope = sqrtf(fmaf(nvy, nvy, nvx * nvx)) / nvz;
    const float curvature = curv_num / pixel;
    return fmaxf(clampf(4.0f * curvature, 0.0f, 1.5f), clampf(fmaf(2.0f, slope, -0.5f), 0.1f, 4.0f));
}

/* NOISE: 0 = none (rs == 0), 1 = plain |rs|*n, 2 = rs < 0 (zf -= rs*n), 3 = slope/curvature modulated */
enum { NZ_NONE = 0, NZ_PLAIN = 1, NZ_NEG = 2, NZ_SLOPE = 3 };
/* RESID: 0 = none, 1 = float pool, 2 = int16 pool */

/* ------------------------------------------------------------------------
 * Specialised kernel: compile-time geometry (TW = tile width, TG = grid
 * divisor, TG even).  Same arithmetic as the generic kernel, far fewer
 * instructions around it (the kernel is issue-bound, profiles/):
 *   - all index maths folds to immediates; the item -> (row pair, quad) split
 *     is a multiply-shift
 *   - the tile-uniform switches (noise variant) are hoisted out of the loop:
 *     one loop instance per variant
 *   - divisions by the tile's pixel size cost 3 FFMA each (pl_fpexact.cuh: the
 *     reciprocal is refined once per thread), sqrt is the 5-instruction IEEE
 *     fast path, no range-check branches
 *   - with TG even the two texels of a quad in x (and in y) share their coarse
 *     lattice indices: zc needs 2 lattice reads per quad, not 8; only texel
 *     (x0, y0) can satisfy the diagonal-flip test
 *   - rows 0..TW-2 are processed as TW/2 full row pairs without bounds checks,
 *     row TW-1 by a short tail; the pad columns of a row (x >= TW) are
 *     "don't care": they are computed like texels (all reads stay inside the
 *     staged window / padded planes) and stored, so sectors are written whole,
 *     and nobody ever reads them (TMA zero-fills x >= TW, downloads skip them)
 * ------------------------------------------------------------------------ */
template <int TW, int TG>
struct Geo {
    static constexpr int W = TW, G = TG;
    static constexpr int PITCH = (TW + 3) & ~3;
    static constexpr int PLANE = TW * PITCH;
    static constexpr int BOX_H = (TW - 5) / 2 + 6;
    static constexpr int BOX_W = (BOX_H + 3) & ~3;
    static constexpr int NK = (TW - 3 + TG) / (2 * TG) + 2;
    static constexpr int QW = PITCH / 2;          /* quads per row pair (incl. pad quads) */
    static constexpr int QH = (TW - 1) / 2;       /* full row pairs */
    static_assert(TW % 2 == 1 && TG % 2 == 0, "odd tile width, even grid divisor");
};

struct QuadCtx {
    const float *winA;         /* staged parent zf window */
    const float *winB;         /* the same window one texel to the right: winB[k] = winA[k + 1] */
    const float *lat;          /* parent zm lattice */
    const int *lutx, *luty;    /* per quad column / row pair: lattice indices + flip bit */
    float *out;                /* zf plane of the output slot */
    float *zms;                /* ZS: a copy of the zm plane in shared memory (same pitch) for the normal pass */
    const __half *nplane;      /* rotated noise plane (columns permuted: pl_noise_col) */
    const void *resid;         /* residual tile origin (slot base + window origin) or NULL */
    int noise_pitch, resid_pitch;
    float rs, ars, pixel, nvz, rcp_pixel, rcp_nvz, resid_scale;
    float zm_floor;            /* 0, or -inf under NO_CLAMP */
    bool has_parent, has_resid, flip, want_stats;
};

/* max(clamp(a, 0, 1.5), clamp(b, 0.1, 4)) = max3(min(a, 1.5), min(b, 4), 0.1) for finite a, b: the lower
 * clamp of a is absorbed by b's (>= 0.1) and clamp(b, 0.1, 4) = max(min(b, 4), 0.1); one FMNMX3 */
__device__ __forceinline__ float amp_clamp(float a, float b)
{
    return fmaxf(fmaxf(fminf(a, 1.5f), fminf(b, 4.0f)), 0.1f);
}
/* noise amplitude factor of two texels: upsampleShader.glsl:166-170 */
__device__ __forceinline__ plf2::F2 amp_of2(const QuadCtx &k, plf2::F2 sx, plf2::F2 sy, plf2::F2 cv)
{
    using namespace plf2;
    const F2 slope = div_rn2(sqrt_rn2(fma2(sy, sy, mul2(sx, sx))), bc(k.nvz), bc(k.rcp_nvz));
    const F2 curvature = div_rn2(cv, bc(k.pixel), bc(k.rcp_pixel));
    const F2 a = mul2(bc(4.0f), curvature), b = fma2(bc(2.0f), slope, bc(-0.5f));
    return make_float2(amp_clamp(a.x, b.x), amp_clamp(a.y, b.y));
}

/* Two horizontally adjacent 2x2 quads (texel columns 4*ip .. 4*ip+3, rows 2*j, 2*j+1) per call; every
 * fp32 operation of the canonical order is issued ONCE for both quads as a packed fp32x2 instruction:
 * lane .x = quad A (columns 4ip, 4ip+1), lane .y = quad B (columns 4ip+2, 4ip+3).
 * TAIL: only the first row of the quads exists (row TW-1). */
/* two adjacent floats of shared memory that do NOT start on an 8-byte boundary, as one register pair: two 4-byte loads
 * straight into the halves of the pair, at a constant byte offset from one shared-space address per quad pair.  Written
 * as PTX so that the compiler does not reuse the halves of neighbouring 8-byte loads instead (it then rebuilds the pair
 * with MOVs: + 5 % instructions, measured) */
template <int OFF>
__device__ __forceinline__ plf2::F2 lds_pair_at(const uint32_t a)
{
    float x, y;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(x) : "r"(a), "n"(OFF));
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(y) : "r"(a), "n"(OFF + 4));
    return make_float2(x, y);
}

template <class GEO, int NZ, int RESID, bool TAIL, bool ZS, bool WB = true>
__device__ __forceinline__ void do_quad2(const QuadCtx &k, const int ip, const int j, float &lo, float &hi)
{
    using namespace plf2;
    constexpr int BW = GEO::BOX_W, PITCH = GEO::PITCH, PLANE = GEO::PLANE;
    const int i = 2 * ip, x0 = 4 * ip, y0 = 2 * j;

    /* residual texels (residualOSH.w * texel): r<x><y><quad> */
    float r00A = 0.0f, r10A = 0.0f, r00B = 0.0f, r10B = 0.0f, r01A = 0.0f, r11A = 0.0f, r01B = 0.0f, r11B = 0.0f;
    if (RESID != 0 && k.has_resid) {
        const int off = y0 * k.resid_pitch + x0;
        if (RESID == 1) {
            const float *rp = static_cast<const float *>(k.resid) + off;
            const float4 t0 = __ldg(reinterpret_cast<const float4 *>(rp));
            r00A = t0.x; r10A = t0.y; r00B = t0.z; r10B = t0.w;
            if (!TAIL) {
                const float4 t1 = __ldg(reinterpret_cast<const float4 *>(rp + k.resid_pitch));
                r01A = t1.x; r11A = t1.y; r01B = t1.z; r11B = t1.w;
            }
        } else {
            /* int16 -> float without the conversion unit: flipping the sign bit makes the value an unsigned
             * s + 32768, PRMT plants its two bytes in the mantissa of 2^23 (0x4B00hhll), one FADD removes
             * 2^23 + 32768: exactly (float) s */
            const short *rp = static_cast<const short *>(k.resid) + off;
            const uint2 t0 = __ldg(reinterpret_cast<const uint2 *>(rp));
            const unsigned int a0 = t0.x ^ 0x80008000u, b0 = t0.y ^ 0x80008000u;
            constexpr float kBias = 8388608.0f + 32768.0f;
            r00A = (__uint_as_float(__byte_perm(a0, 0x4B000000u, 0x7410u)) - kBias) * k.resid_scale;
            r10A = (__uint_as_float(__byte_perm(a0, 0x4B000000u, 0x7432u)) - kBias) * k.resid_scale;
            r00B = (__uint_as_float(__byte_perm(b0, 0x4B000000u, 0x7410u)) - kBias) * k.resid_scale;
            r10B = (__uint_as_float(__byte_perm(b0, 0x4B000000u, 0x7432u)) - kBias) * k.resid_scale;
            if (!TAIL) {
                const uint2 t1 = __ldg(reinterpret_cast<const uint2 *>(rp + k.resid_pitch));
                const unsigned int a1 = t1.x ^ 0x80008000u, b1 = t1.y ^ 0x80008000u;
                r01A = (__uint_as_float(__byte_perm(a1, 0x4B000000u, 0x7410u)) - kBias) * k.resid_scale;
                r11A = (__uint_as_float(__byte_perm(a1, 0x4B000000u, 0x7432u)) - kBias) * k.resid_scale;
                r01B = (__uint_as_float(__byte_perm(b1, 0x4B000000u, 0x7410u)) - kBias) * k.resid_scale;
                r11B = (__uint_as_float(__byte_perm(b1, 0x4B000000u, 0x7432u)) - kBias) * k.resid_scale;
            }
        }
    }

    /* noise texels (already rotated); one 8-byte load = (even texel of A, of B), (odd texel of A, of B) */
    F2 n00 = bc(0.0f), n10 = bc(0.0f), n01 = bc(0.0f), n11 = bc(0.0f);
    if (NZ != NZ_NONE) {
        const __half *np = k.nplane + y0 * k.noise_pitch + x0;
        const uint2 t0 = __ldg(reinterpret_cast<const uint2 *>(np));
        n00 = __half22float2(*reinterpret_cast<const __half2 *>(&t0.x));
        n10 = __half22float2(*reinterpret_cast<const __half2 *>(&t0.y));
        if (!TAIL) {
            const uint2 t1 = __ldg(reinterpret_cast<const uint2 *>(np + k.noise_pitch));
            n01 = __half22float2(*reinterpret_cast<const __half2 *>(&t1.x));
            n11 = __half22float2(*reinterpret_cast<const __half2 *>(&t1.y));
        }
    }

    /* cz[c][r] = parent.zf[bx + r, by + c] -> z<c><r> = (quad A, quad B): quad B's taps are quad A's one
     * window column further right, so even r are aligned pairs of winA, odd r aligned pairs of winB */
    /* WB = false (the slim layout of the fused kernel, pl_pair.cu): no second copy of the window; the odd pairs are two
     * 4-byte loads into one register pair instead of one 8-byte load -- 8 more LDS per quad pair, 12 KB less shared memory */
    const float *wa = k.winA + j * BW + i, *wb = WB ? k.winB + j * BW + i : wa + 1;
    const uint32_t sa = WB ? 0u : smem_u32(wa);
#define LDP(p) (*reinterpret_cast<const F2 *>(p))
#define LDO(row, col) (WB ? LDP(wb + (row) * BW + (col)) : lds_pair_at<4 * ((row) * BW + (col) + 1)>(sa))
    const F2 z00 = LDP(wa), z01 = LDO(0, 0), z02 = LDP(wa + 2), z03 = LDO(0, 2);
    const F2 z10 = LDP(wa + BW), z11 = LDO(1, 0), z12 = LDP(wa + BW + 2), z13 = LDO(1, 2);
    const F2 z20 = LDP(wa + 2 * BW), z21 = LDO(2, 0), z22 = LDP(wa + 2 * BW + 2), z23 = LDO(2, 2);
    F2 z30 = bc(0.0f), z31 = bc(0.0f), z32 = bc(0.0f), z33 = bc(0.0f);
    if (!TAIL) { z30 = LDP(wa + 3 * BW); z31 = LDO(3, 0); z32 = LDP(wa + 3 * BW + 2); z33 = LDO(3, 2); }
#undef LDO
#undef LDP

    /* noise term scale per texel: T<x><y> */
    F2 t00 = bc(0.0f), t10 = bc(0.0f), t01 = bc(0.0f), t11 = bc(0.0f);
    if (NZ == NZ_PLAIN) {
        t00 = t10 = t01 = t11 = bc(k.ars);
    } else if (NZ == NZ_NEG) {
        t00 = t10 = t01 = t11 = bc(-k.rs);
    } else if (NZ == NZ_SLOPE) {
        {   /* parity 0: slopex/slopey/curvature matrices [0] */
            const F2 sx = sub2(z10, z12);
            const F2 sy = sub2(z01, z21);
            const F2 cv = sub2(add2(neg(z01), sub2(fma2(z11, bc(4.0f), neg(z10)), z12)), z21);
            t00 = mul2(amp_of2(k, sx, sy, cv), bc(k.rs));
        }
        {   /* parity 1: [1] */
            const F2 sx = plf2::chain4(z10, z11, z12, z13, 0.5f, 0.5f, -0.5f, -0.5f);
            const F2 lowr = fma2(z22, bc(-0.5f), mul2(z21, bc(-0.5f)));
            const F2 sy = add2(fma2(z02, bc(0.5f), mul2(z01, bc(0.5f))), lowr);
            const F2 cv = add2(add2(fma2(z02, bc(-0.5f), mul2(z01, bc(-0.5f))), plf2::chain4(z10, z11, z12, z13, -0.5f, 1.5f, 1.5f, -0.5f)), lowr);
            t10 = mul2(amp_of2(k, sx, sy, cv), bc(k.rs));
        }
        if (!TAIL) {
            {   /* parity 2: [2]; every bare product below has a power-of-two weight (exact) */
                const F2 sx = add2(fma2(z12, bc(-0.5f), mul2(z10, bc(0.5f))), fma2(z22, bc(-0.5f), mul2(z20, bc(0.5f))));
                const F2 sy = add2(add2(add2(mul2(z01, bc(0.5f)), mul2(z11, bc(0.5f))), mul2(z21, bc(-0.5f))), mul2(z31, bc(-0.5f)));
                const F2 cv = add2(add2(add2(mul2(z01, bc(-0.5f)), fma2(z12, bc(-0.5f), fma2(z11, bc(1.5f), mul2(z10, bc(-0.5f))))),
                                        fma2(z22, bc(-0.5f), fma2(z21, bc(1.5f), mul2(z20, bc(-0.5f))))), mul2(z31, bc(-0.5f)));
                t01 = mul2(amp_of2(k, sx, sy, cv), bc(k.rs));
            }
            {   /* parity 3: [3] */
                const F2 sx = add2(plf2::chain4(z10, z11, z12, z13, 0.25f, 0.25f, -0.25f, -0.25f),
                                   plf2::chain4(z20, z21, z22, z23, 0.25f, 0.25f, -0.25f, -0.25f));
                const F2 lowr = fma2(z32, bc(-0.25f), mul2(z31, bc(-0.25f)));
                const F2 sy = add2(add2(add2(fma2(z02, bc(0.25f), mul2(z01, bc(0.25f))), fma2(z12, bc(0.25f), mul2(z11, bc(0.25f)))),
                                        fma2(z22, bc(-0.25f), mul2(z21, bc(-0.25f)))), lowr);
                const F2 cv = add2(add2(add2(fma2(z02, bc(-0.25f), mul2(z01, bc(-0.25f))), plf2::chain4(z10, z11, z12, z13, -0.25f, 0.5f, 0.5f, -0.25f)),
                                        plf2::chain4(z20, z21, z22, z23, -0.25f, 0.5f, 0.5f, -0.25f)), lowr);
                t11 = mul2(amp_of2(k, sx, sy, cv), bc(k.rs));
            }
        }
    }

    /* upsampleMatrix[0..3]; at level 0 the window is all zeros and the term adds 0 */
    const float W1 = -1.0f / 16.0f, W9 = 9.0f / 16.0f;
    const float V1 = 1.0f / 256.0f, V9 = -9.0f / 256.0f, V81 = 81.0f / 256.0f;
    const F2 u00 = z11;
    const F2 u10 = plf2::chain4(z10, z11, z12, z13, W1, W9, W9, W1);
    F2 u01 = bc(0.0f), u11 = bc(0.0f);
    if (!TAIL) {
        /* ((z01*W1 + z11*W9) + z21*W9) + z31*W1: products by W1 are exact, so they may ride in an fma;
         * the product z21*W9 must round on its own and ptxas would contract a packed multiply into the
         * packed add that follows (pl_f2.cuh) -> that one step is scalar */
        const F2 x = fma2(z01, bc(W1), mul2(z11, bc(W9)));
        const float qx = __fmul_rn(z21.x, W9), qy = __fmul_rn(z21.y, W9);
        u01 = fma2(z31, bc(W1), make_float2(__fadd_rn(x.x, qx), __fadd_rn(x.y, qy)));
        u11 = add2(add2(add2(plf2::chain4(z00, z01, z02, z03, V1, V9, V9, V1), plf2::chain4(z10, z11, z12, z13, V9, V81, V81, V9)),
                        plf2::chain4(z20, z21, z22, z23, V9, V81, V81, V9)), plf2::chain4(z30, z31, z32, z33, V1, V9, V9, V1));
    }

    /* zf = residual + scale * noise, then + upsample.  The last step is scalar so that its results can
     * land directly in store order (A.x0, A.x0+1, B.x0+2, B.x0+3): one 16-byte store per row and plane */
    float f00A, f10A, f00B, f10B, f01A = 0.0f, f11A = 0.0f, f01B = 0.0f, f11B = 0.0f;
    if (NZ == NZ_NONE) {
        f00A = r00A; f10A = r10A; f00B = r00B; f10B = r10B;
        f01A = r01A; f11A = r11A; f01B = r01B; f11B = r11B;
    } else {
        f00A = fmaf(t00.x, n00.x, r00A); f00B = fmaf(t00.y, n00.y, r00B);
        f10A = fmaf(t10.x, n10.x, r10A); f10B = fmaf(t10.y, n10.y, r10B);
        if (!TAIL) {
            f01A = fmaf(t01.x, n01.x, r01A); f01B = fmaf(t01.y, n01.y, r01B);
            f11A = fmaf(t11.x, n11.x, r11A); f11B = fmaf(t11.y, n11.y, r11B);
        }
    }
    /* level 0: zc = zf before the (zero) upsample term: keep those for the zc plane */
    const float4 g0 = make_float4(f00A, f10A, f00B, f10B), g1 = make_float4(f01A, f11A, f01B, f11B);
    f00A = f00A + u00.x; f00B = f00B + u00.y; f10A = f10A + u10.x; f10B = f10B + u10.y;
    if (!TAIL) { f01A = f01A + u01.x; f01B = f01B + u01.y; f11A = f11A + u11.x; f11B = f11B + u11.y; }

    /* zm = max(zf, 0), or zf itself under NO_CLAMP (floor = -inf) */
    const float m00A = fmaxf(f00A, k.zm_floor), m10A = fmaxf(f10A, k.zm_floor), m00B = fmaxf(f00B, k.zm_floor), m10B = fmaxf(f10B, k.zm_floor);
    float m01A = 0.0f, m11A = 0.0f, m01B = 0.0f, m11B = 0.0f;
    if (!TAIL) { m01A = fmaxf(f01A, k.zm_floor); m11A = fmaxf(f11A, k.zm_floor); m01B = fmaxf(f01B, k.zm_floor); m11B = fmaxf(f11B, k.zm_floor); }

    float *o0 = k.out + y0 * PITCH + x0;
    *reinterpret_cast<float4 *>(o0) = make_float4(f00A, f10A, f00B, f10B);
    *reinterpret_cast<float4 *>(o0 + 2 * PLANE) = make_float4(m00A, m10A, m00B, m10B);
    if (ZS) *reinterpret_cast<float4 *>(k.zms + y0 * PITCH + x0) = make_float4(m00A, m10A, m00B, m10B);
    if (!TAIL) {
        *reinterpret_cast<float4 *>(o0 + PITCH) = make_float4(f01A, f11A, f01B, f11B);
        *reinterpret_cast<float4 *>(o0 + 2 * PLANE + PITCH) = make_float4(m01A, m11A, m01B, m11B);
        if (ZS) *reinterpret_cast<float4 *>(k.zms + (y0 + 1) * PITCH + x0) = make_float4(m01A, m11A, m01B, m11B);
    }

    if (k.has_parent) {
        /* TG even: both texels of a quad in x (in y) round to the same lattice column (row);
         * zc1 = zm[round_x, floor_y], zc3 = zm[floor_x, round_y] */
        const int2 lx = *reinterpret_cast<const int2 *>(k.lutx + i);
        const int ly = k.luty[j];
        const int kfy = ly & 0x7fff, kry = (ly >> 16) & 0x7fff;      /* premultiplied by NK */
        const int krxA = lx.x & 0x7fff, kfxA = (lx.x >> 16) & 0x7fff;
        const int krxB = lx.y & 0x7fff, kfxB = (lx.y >> 16) & 0x7fff;
        const float zc1A = k.lat[krxA + kfy], zc3A = k.lat[kfxA + kry];
        const float zc1B = k.lat[krxB + kfy], zc3B = k.lat[kfxB + kry];
        const float zcA = (zc1A + zc3A) * 0.5f, zcB = (zc1B + zc3B) * 0.5f;
        if (!TAIL) *reinterpret_cast<float4 *>(o0 + PLANE + PITCH) = make_float4(zcA, zcA, zcB, zcB);
        float c00A = zcA, c00B = zcB;
        if (k.flip && (ly & 0x8000)) {   /* only texel (x0, y0) of a quad can sit on a flipped diagonal */
            if (lx.x & 0x8000) {
                const float zc0 = k.lat[kfxA + kfy], zc2 = k.lat[krxA + kry];
                c00A = (zc3A + zc1A >= zc0 + zc2 ? zc1A + zc3A : zc0 + zc2) * 0.5f;
            }
            if (lx.y & 0x8000) {
                const float zc0 = k.lat[kfxB + kfy], zc2 = k.lat[krxB + kry];
                c00B = (zc3B + zc1B >= zc0 + zc2 ? zc1B + zc3B : zc0 + zc2) * 0.5f;
            }
        }
        *reinterpret_cast<float4 *>(o0 + PLANE) = make_float4(c00A, zcA, c00B, zcB);
    } else {
        *reinterpret_cast<float4 *>(o0 + PLANE) = g0;
        if (!TAIL) *reinterpret_cast<float4 *>(o0 + PLANE + PITCH) = g1;
    }

    if (k.want_stats && !TAIL) {   /* TileSamplerZ.cpp:60-64: texels [2, W-3]^2 of zm; row W-1 is outside */
        const bool pA = x0 >= 2 && x0 + 1 <= GEO::W - 3;             /* both columns of quad A */
        const bool pBe = x0 + 2 <= GEO::W - 3, pBo = x0 + 3 <= GEO::W - 3;
        const bool py0 = j >= 1, py1 = j >= 1 && y0 + 1 <= GEO::W - 3;   /* y0 <= W-3 always holds here */
        /* quad B's odd column is outside only in the last real quad pair: fold it onto the even column */
        const float b0 = pBo ? m10B : m00B, b1 = pBo ? m11B : m01B;
        if (pA && py0) { lo = fminf(fminf(lo, m00A), m10A); hi = fmaxf(fmaxf(hi, m00A), m10A); }
        if (pA && py1) { lo = fminf(fminf(lo, m01A), m11A); hi = fmaxf(fmaxf(hi, m01A), m11A); }
        if (pBe && py0) { lo = fminf(fminf(lo, m00B), b0); hi = fmaxf(fmaxf(hi, m00B), b0); }
        if (pBe && py1) { lo = fminf(fminf(lo, m01B), b1); hi = fmaxf(fmaxf(hi, m01B), b1); }
    }
}

template <class GEO, int NZ, int RESID, int NT, bool ZS, bool WB = true>
__device__ __forceinline__ void tile_loop(const QuadCtx &k, const int tid, float &lo, float &hi)
{
    constexpr int QP = GEO::QW / 2, QH = GEO::QH;   /* quad pairs per row pair, full row pairs */
    constexpr int DJ = NT / QP, DI = NT - DJ * QP;   /* one stride of NT items in (row pair, quad pair) */
    int j = tid / QP, ip = tid - j * QP;
    while (j < QH) {
        do_quad2<GEO, NZ, RESID, false, ZS, WB>(k, ip, j, lo, hi);
        ip += DI; j += DJ;
        if (ip >= QP) { ip -= QP; j += 1; }
    }
    if (tid < QP) do_quad2<GEO, NZ, RESID, true, ZS, WB>(k, tid, QH, lo, hi);
}

/* Shared memory one tile needs (all threads of the CTA pass the same pointers). */
template <int TW, int TG, bool WB = true>
struct ElevSmem {
    using GEO = Geo<TW, TG>;
    static constexpr int WIN = GEO::BOX_H * GEO::BOX_W;          /* floats per window copy */
    static constexpr int LAT = GEO::NK * GEO::NK;
    static constexpr int FLOATS = (WB ? 2 : 1) * WIN + ((LAT + 3) & ~3) + 2 * GEO::QW + 2 * 32;   /* window(s), lattice, luts, reduction */
    static_assert(WIN % 4 == 0, "the second window copy is written as float4");
};

/* One tile: stage the parent window (TMA), build the tables, run the quad loop, reduce the statistics.
 * smem: ElevSmem<TW,TG>::FLOATS floats, 128-byte aligned; bar: an mbarrier initialised with count 1 whose
 * current phase parity is `parity`.  All NT threads of the CTA call this; it contains __syncthreads(). */
/* defer_stats: the caller has a barrier of its own right after this call: the per-warp statistics are left in shared
 * memory and the caller finishes them with elevation_stats_finish after that barrier (one barrier less per tile) */
template <int TW, int TG, int NT, bool WB = true>
__device__ __forceinline__ void elevation_stats_finish(const ElevArgs &a, const pl_elev_req &rq, const float *smem)
{
    using SM = ElevSmem<TW, TG, WB>;
    const float *red_lo = smem + SM::FLOATS - 64, *red_hi = red_lo + 32;
    float lo = red_lo[0], hi = red_hi[0];
#pragma unroll
    for (int q = 1; q < NT / 32; ++q) { lo = fminf(lo, red_lo[q]); hi = fmaxf(hi, red_hi[q]); }
    a.stats[rq.out_slot] = make_float2(lo, hi);
}

template <int TW, int TG, int RESID, int NT, bool ZS, bool WB = true>
__device__ __forceinline__ void elevation_tile(const CUtensorMap *tm, const ElevArgs &a, const pl_elev_req &rq, float *smem,
                                               uint64_t *bar, const uint32_t parity, float *zms, const int tid,
                                               const bool defer_stats = false)
{
    using GEO = Geo<TW, TG>;
    using SM = ElevSmem<TW, TG, WB>;
    constexpr int W = GEO::W, G = GEO::G, NK = GEO::NK, QW = GEO::QW;
    static_assert(QW % 2 == 0 && GEO::BOX_W % 4 == 0 && GEO::BOX_W >= QW + 4, "quad pairs read aligned float pairs inside a window row");
    float *winA = smem;
    float *winB = winA + SM::WIN;                       /* WB only */
    float *lat = winA + (WB ? 2 : 1) * SM::WIN;
    int *lutx = reinterpret_cast<int *>(lat + ((SM::LAT + 3) & ~3));
    int *luty = lutx + QW;
    float *red_lo = reinterpret_cast<float *>(luty + QW), *red_hi = red_lo + 32;   /* = smem + SM::FLOATS - 64 (elevation_stats_finish) */

    const bool has_parent = rq.parent_slot >= 0;

    if (tid == 0 && has_parent) {
        mbar_expect_tx(bar, (uint32_t) (GEO::BOX_W * GEO::BOX_H * sizeof(float)));
        tma_load_3d(winA, tm, bar, rq.dx, rq.dy, rq.parent_slot * 3);
    }
    /* lattice indices of quad column / row pair q (texel 2q): round and floor variants */
    if (tid < QW) {
        const int ij = 2 * tid - 2;
        const int kr = floordiv(ij + G, 2 * G) + 1, kf = floordiv(ij, 2 * G) + 1;
        int m = ij % (2 * G);
        if (m < 0) m += 2 * G;
        const int fl = (m == G) ? 0x8000 : 0;
        lutx[tid] = kr | fl | (kf << 16);
        luty[tid] = (kf * NK) | fl | ((kr * NK) << 16);
    }
    if (has_parent) {
        const float *pzm = a.elev + (size_t) rq.parent_slot * 3 * GEO::PLANE + 2 * GEO::PLANE;
        for (int q = tid; q < NK * NK; q += NT) {
            const int kx = q % NK - 1, ky = q / NK - 1;
            const int px = min(max(2 + G * kx + rq.dx, 0), W - 1);   /* CLAMP_TO_EDGE */
            const int py = min(max(2 + G * ky + rq.dy, 0), W - 1);
            lat[q] = __ldcg(pzm + py * GEO::PITCH + px);   /* L2: the parent may have been finished by another CTA of this launch */
        }
    } else {
        for (int q = tid; q < GEO::BOX_W * GEO::BOX_H; q += NT) { winA[q] = 0.0f; if (WB) winB[q] = 0.0f; }
        for (int q = tid; q < NK * NK; q += NT) lat[q] = 0.0f;
    }
    __syncthreads();
    if (has_parent && !WB) mbar_wait(bar, parity);
    if (has_parent && WB) {
        mbar_wait(bar, parity);
        /* the window a second time, one texel to the right: packed loads of odd window columns become
         * aligned pairs of the copy.  (A second TMA load at dx + 1 is not possible: without swizzle the
         * inner coordinate of a tiled fp32 copy must start on a 16-byte boundary -- measured on B200.) */
        constexpr int G4 = GEO::BOX_W / 4;
        for (int q = tid; q < GEO::BOX_H * G4; q += NT) {
            const int g = q % G4;
            const float4 v = *reinterpret_cast<const float4 *>(winA + 4 * q);
            const float nxt = g + 1 < G4 ? winA[4 * q + 4] : 0.0f;
            *reinterpret_cast<float4 *>(winB + 4 * q) = make_float4(v.y, v.z, v.w, nxt);
        }
        __syncthreads();
    }

    QuadCtx k;
    k.winA = winA;
    k.winB = WB ? winB : winA;
    k.lat = lat;
    k.lutx = lutx;
    k.luty = luty;
    k.out = a.elev + (size_t) rq.out_slot * 3 * GEO::PLANE;
    k.zms = zms;
    k.nplane = a.noise + (size_t) (rq.noise_r * 6 + rq.noise_l) * a.noise_plane;
    k.noise_pitch = a.noise_pitch;
    k.resid_pitch = a.resid_pitch;
    k.has_resid = RESID != 0 && rq.resid_slot >= 0;
    k.resid = nullptr;
    if (k.has_resid) {
        const size_t off = (size_t) rq.resid_slot * a.resid_slot_elems + (size_t) rq.ry * a.resid_pitch + rq.rx;
        k.resid = RESID == 1 ? static_cast<const void *>(static_cast<const float *>(a.resid) + off)
                             : static_cast<const void *>(static_cast<const short *>(a.resid) + off);
    }
    k.rs = rq.rs;
    k.ars = fabsf(rq.rs);
    k.pixel = rq.pixel_size;
    k.nvz = 2.0f * rq.pixel_size;
    k.rcp_pixel = plfp::rcp_rn(k.pixel);
    k.rcp_nvz = plfp::rcp_rn(k.nvz);
    k.resid_scale = a.resid_scale;
    k.has_parent = has_parent;
    k.flip = a.flip != 0;
    k.zm_floor = a.no_clamp ? -INFINITY : 0.0f;
    k.want_stats = a.want_stats != 0;

    float lo = INFINITY, hi = -INFINITY;
    /* tile-uniform noise variant; rs == 0 adds exactly 0 in every variant */
    const int nz = rq.rs == 0.0f ? NZ_NONE : (a.noise_mode == PL_NOISE_PLAIN ? NZ_PLAIN : (rq.rs < 0.0f ? NZ_NEG : NZ_SLOPE));
    switch (nz) {
    case NZ_NONE: tile_loop<GEO, NZ_NONE, RESID, NT, ZS, WB>(k, tid, lo, hi); break;
    case NZ_PLAIN: tile_loop<GEO, NZ_PLAIN, RESID, NT, ZS, WB>(k, tid, lo, hi); break;
    case NZ_NEG: tile_loop<GEO, NZ_NEG, RESID, NT, ZS, WB>(k, tid, lo, hi); break;
    default: tile_loop<GEO, NZ_SLOPE, RESID, NT, ZS, WB>(k, tid, lo, hi); break;
    }

    if (k.want_stats) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, s));
            hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, s));
        }
        if ((tid & 31) == 0) { red_lo[tid >> 5] = lo; red_hi[tid >> 5] = hi; }
        if (defer_stats) return;
        __syncthreads();
        if (tid == 0) {
#pragma unroll
            for (int q = 1; q < NT / 32; ++q) { lo = fminf(lo, red_lo[q]); hi = fmaxf(hi, red_hi[q]); }
            a.stats[rq.out_slot] = make_float2(lo, hi);
        }
    }
}


}  // namespace plelev
#endif
