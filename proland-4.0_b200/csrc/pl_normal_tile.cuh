/*
 * pl_normal_tile.cuh -- device code of the NormalProducer pass for ONE tile of the shipped geometry
 * (compile-time tile width, border 2, RG8), shared by the normal kernel (pl_normal.cu) and the fused
 * elevation+normal kernel (pl_pair.cu).
 *
 * Reference: normalShader.glsl:60-125 after NormalProducer::doCreateTile (NormalProducer.cpp:164-289)
 * has set the uniforms.  Arithmetic: canonical order of oracle/orc_fp.h.
 */
#ifndef PL_NORMAL_TILE_CUH
#define PL_NORMAL_TILE_CUH

#include "pl_internal.h"
#include "pl_fpexact.cuh"
#include "pl_f2.cuh"

namespace plnorm {


struct NormArgs {
    const float *elev;       /* elevation pool base */
    uint8_t *norm;           /* normal pool base */
    const pl_norm_req *reqs;
    int W;                   /* normal tile width */
    int EW, epitch, eplane;  /* elevation tile width, row pitch, plane elems */
    int border;
    int linear;              /* elevation storage filter */
    int sphere;
    int channels;            /* 2 (RG8) or 4 (RGBA8: fine + coarse normal) */
    int grid;                /* tileSDF.y */
    int parent_linear;       /* normal storage filter (parent coarse normal fetch) */
    int nbands, max_rows;    /* bands per tile, rows of the largest band */
    int fast;                /* pl_norm_scene.arith == PL_ARITH_FAST */
    long long norm_slot_bytes;
    /* push of finished normal tiles to the peer GPUs (pl_pool_attach_peers): byte distance from this GPU's
     * normal pool to each peer's mapping of its own */
    int npeers;
    long long peer_delta[7];
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
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void *dst, const void *src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

__device__ __forceinline__ float dot3(float a0, float a1, float a2, float b0, float b1, float b2)
{
    return fmaf(a2, b2, fmaf(a1, b1, a0 * b0));
}
__device__ __forceinline__ float dot4(const float *m, float v0, float v1, float v2, float v3)
{
    return fmaf(m[3], v3, fmaf(m[2], v2, fmaf(m[1], v1, m[0] * v0)));
}
/* OpenGL 3.3 spec 2.1.5: float -> unorm8, round to nearest */
__device__ __forceinline__ unsigned int unorm8(float f)
{
    if (!(f > 0.0f)) return 0u;
    if (f >= 1.0f) return 255u;
    return (unsigned int) __float2int_rn(f * 255.0f);
}
__device__ __forceinline__ int floordiv(int a, int b) { int q = a / b; return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q; }

/* .xy of the parent's RGBA8 normal tile at texel coordinate (cx, cy) (already offset by +0.25,
 * NormalProducer.cpp:201-205) through the normal storage's filter; unorm8 -> float is c / 255
 * (OpenGL 3.3 spec 2.1.5), CLAMP_TO_EDGE */
__device__ __forceinline__ float2 fetch_parent_xy(const uchar4 *parent, int W, bool linear, float cx, float cy)
{
    auto PN = [&](int i, int j) {
        const uchar4 t = __ldg(parent + min(max(j, 0), W - 1) * W + min(max(i, 0), W - 1));
        return make_float2((float) t.x / 255.0f, (float) t.y / 255.0f);
    };
    if (!linear) return PN((int) floorf(cx), (int) floorf(cy));
    const float fx = cx - 0.5f, fy = cy - 0.5f;
    const int i0 = (int) floorf(fx), j0 = (int) floorf(fy);
    const float fa = fx - (float) i0, fb = fy - (float) j0;
    const float2 t00 = PN(i0, j0), t10 = PN(i0 + 1, j0), t01 = PN(i0, j0 + 1), t11 = PN(i0 + 1, j0 + 1);
    const float w11 = fa * fb, w01 = (1.0f - fa) * fb, w10 = fa * (1.0f - fb), w00 = (1.0f - fa) * (1.0f - fb);
    return make_float2(fmaf(w11, t11.x, fmaf(w01, t01.x, fmaf(w10, t10.x, w00 * t00.x))),
                       fmaf(w11, t11.y, fmaf(w01, t01.y, fmaf(w10, t10.y, w00 * t00.y))));
}

/* ------------------------------------------------------------------------
 * Specialised kernel: compile-time geometry (TW normal tile width, border 2).
 * Same arithmetic as the generic kernel; the pass is issue-bound (profiles/),
 * so everything here is about the instruction count per texel:
 *   - ONE CTA PER TILE: the whole zm plane of the tile arrives by one bulk copy
 *     (its rows are contiguous in the pitched plane); the CTA then walks the tile
 *     in bands of kTileBand rows, so the per-CTA set-up (request, uv table,
 *     barrier) is paid once per tile, and the two grid rows a band shares with
 *     the next one are carried over instead of recomputed
 *   - every thread works on a PAIR of horizontally adjacent grid points and on
 *     a 2x2 block of texels; all fp32 maths is packed FFMA2/FMUL2/FADD2
 *     (pl_f2.cuh): half the issue slots for the same IEEE results
 *   - index maths folds to immediates / multiply-shifts
 *   - the four quotients alpha*L/dot(alpha,L) share ONE refined reciprocal
 *     (pl_fpexact.cuh: 3 FFMA per IEEE quotient), normalisation is the
 *     5-instruction IEEE sqrt + 3-instruction IEEE reciprocal, no branches
 *   - positions live in shared memory as three planes of row pitch GWP (even),
 *     rows shifted so that every access of the normal phase is an aligned pair
 *   - the RG8 texels go from registers straight to HBM (2-byte stores; the L2
 *     merges the two halves of every sector before it is written back)
 *   - SPHERE / LINEAR are template parameters: no per-point tests
 * ------------------------------------------------------------------------ */
constexpr int kTileBand = 20;     /* texel rows per band: 20 x 50 point pairs and 10 x 49 texel blocks fill 256 threads 4x and 2x */

template <int TW>
struct NGeo {
    static constexpr int W = TW, B = 2;
    static constexpr int EW = TW + 2 * B;
    static constexpr int EPITCH = (EW + 3) & ~3;
    static constexpr int EPLANE = EW * EPITCH;
    static constexpr int GW = TW + 2;                 /* grid points X = -1 .. W */
    static constexpr int GWP = ((GW + 1) & ~1) + 4;   /* row pitch of a position plane: even, room for the row shift and the pad pairs */
    static constexpr int GPAIRS = (GW + 2) / 2;       /* grid-point pairs per row (covers an odd start) */
    static constexpr int XPAIRS = (TW + 2) / 2;       /* texel pairs per row (covers an odd start) */
    static constexpr int NBANDS = (TW + kTileBand - 1) / kTileBand;
    static constexpr int LAST_ROWS = TW - (NBANDS - 1) * kTileBand;
    static constexpr int POS_ROWS = kTileBand + 2;    /* grid rows a band touches */
    static constexpr int POS_PLANE = POS_ROWS * GWP;
    static constexpr int ULUT = (GW + 4) & ~1;        /* u of X = -2 .. W+1, twice (second copy one entry further) */
    static constexpr int ROWTAB = POS_ROWS * 16;      /* PL_ARITH_FAST: 16 floats per grid row of a band (row_table) */
    static constexpr size_t SMEM = 16 + (size_t) EPLANE * 4 + (size_t) 3 * POS_PLANE * 4 + (size_t) 2 * ULUT * 4 + (size_t) ROWTAB * 4;   /* 16: guard floats in front of the zm plane */
    static_assert(EPITCH % 2 == 0 && EPITCH >= GW + 3, "paired loads stay inside a staged row");
    static_assert(kTileBand % 4 == 0, "the row-shift pattern restarts with every band");
    static_assert((LAST_ROWS + 1) / 2 * 2 + 2 <= POS_ROWS, "the last band's blocks stay inside the position rows");
    static_assert((2 * GWP * 3) % 4 == 0 && GWP % 4 == 0, "the carried rows move as float4");
};

/* Shared-memory layout of the position planes.  A thread of the normal phase owns a 2x2 block of
 * texels: it needs 4 consecutive grid points of the block's two rows and the 2 middle ones of the
 * rows below / above.  Grid row g is stored shifted right by shift(g) = ((g + 3) >> 1) & 1 elements and
 * texel blocks of block row k start at x = (k & 1) - 1 (mod 2): with that every one of those accesses is
 * an ALIGNED 8-byte pair (shared-memory wavefronts are a limiter of this kernel, profiles/):
 *   block row k even (x even): rows 2k+1, 2k+2 shift 0 -> pairs at x, x+2; rows 2k, 2k+3 shift 1 -> pair at x+1+1
 *   block row k odd  (x odd) : rows 2k+1, 2k+2 shift 1 -> pairs at x+1, x+3; rows 2k, 2k+3 shift 0 -> pair at x+1
 * Bands start on multiples of 4 rows, so the shift of a row depends only on its row inside the band. */
__device__ __forceinline__ int row_shift(int g) { return ((g + 3) >> 1) & 1; }

/* Normals of one tile.  zs: the tile's zm plane (EW rows of pitch EPITCH) in shared memory; pos: 3
 * position planes of POS_ROWS x GWP; ulut: the two uv tables; out: the tile's RG8 texels in HBM.
 * All threads of the CTA call this; it contains __syncthreads(). */
/* PL_ARITH_FAST (pl_norm_scene.arith, template FAST): the tolerance contract of the normal pass.  Elevations stay
 * bit-exact (children are built from them); the normal tile -- a terminal, unorm8-quantised product -- may differ
 * from the canonical evaluation by at most ONE unorm8 step on < 1e-3 of its bytes (tests/test_gpu_fast.py):
 *   - world positions on a sphere (levels whose smoothstep factor is exactly 1, i.e. quad size <= R/64: level >= 7
 *     on a planet; the others keep the exact code): the shader's
 *         alphaPrime = alpha*L / dot(alpha, L);  p = C*alphaPrime + h * (N*alphaPrime)
 *     is p = (C*(alpha*L) + h * N*(alpha*L)) / dot(alpha, L): the three numerators and the denominator are
 *     bilinear in (u, v), so a grid ROW carries their values at the row's two ends (row_table: 7 start values,
 *     7 differences) and a grid point costs 7 fused lerps, ONE reciprocal (MUFU.RCP + one Newton step) shared by
 *     the three components, 3 fma and 3 multiplies -- 15 packed instructions per pair of points instead of 56
 *   - normalisation by MUFU.RSQ, folded with the unorm8 scale into the tangent-frame product:
 *         byte = round(127.5 * rsqrt(n.n) * dot(w2t_row, n) + 127.5)
 *     (no IEEE square root, no reciprocal, no clamp: |t| <= 1 + 1e-6 keeps the magic-add rounding inside 0..255) */
/* the coarse normal of an RGBA8 texel (normalShader.glsl:100-114): the parent tile's normal at the two coarse mesh
 * vertices around (x, y), averaged, unpacked and -- on a sphere -- rotated by parentToTangentFrame; without a parent
 * (normalOSL.x = -1) the fine normal (tx, ty) itself.  -> b | a << 8, the texel's upper two bytes.  Same operations, same
 * order as the runtime-geometry kernel (pl_normal.cu). */
__device__ __forceinline__ unsigned int coarse_normal_ba(const NormArgs &a, const pl_norm_req &rq, const int W, const bool sphere,
                                                         const int x, const int y, const float tx, const float ty)
{
    float ncx = tx, ncy = ty;
    if (rq.parent_slot >= 0) {
        const uchar4 *parent = reinterpret_cast<const uchar4 *>(a.norm + (size_t) rq.parent_slot * a.norm_slot_bytes);
        const int g = a.grid;
        const float offx = (float) rq.ptx * ((float) W / 2.0f) + 0.25f, offy = (float) rq.pty * ((float) W / 2.0f) + 0.25f;
        const float2 nc0 = fetch_parent_xy(parent, W, a.parent_linear != 0, (float) (g * floordiv(x + g, 2 * g)) + offx,
                                           (float) (g * floordiv(y, 2 * g)) + offy);
        const float2 nc1 = fetch_parent_xy(parent, W, a.parent_linear != 0, (float) (g * floordiv(x, 2 * g)) + offx,
                                           (float) (g * floordiv(y + g, 2 * g)) + offy);
        ncx = fmaf((nc0.x + nc1.x) * 0.5f, 2.0f, -1.0f);
        ncy = fmaf((nc0.y + nc1.y) * 0.5f, 2.0f, -1.0f);
        if (sphere) {
            const float ncz = sqrtf(1.0f - fmaf(ncy, ncy, ncx * ncx));
            const float qx = dot3(rq.p2t[0], rq.p2t[1], rq.p2t[2], ncx, ncy, ncz);
            const float qy = dot3(rq.p2t[3], rq.p2t[4], rq.p2t[5], ncx, ncy, ncz);
            ncx = qx;
            ncy = qy;
        }
    }
    return unorm8(fmaf(ncx, 0.5f, 0.5f)) | (unorm8(fmaf(ncy, 0.5f, 0.5f)) << 8);
}

/* C4: an RGBA8 normal storage (tileSDF.z = 1): `out` addresses 4-byte texels (r, g, b, a) = (fine.xy, coarse.xy); the
 * exact arithmetic only (FAST = false) */
template <int TW, bool SPHERE, bool LINEAR, int NT, bool PUSH = false, bool FAST = false, bool C4 = false>
__device__ __forceinline__ void normal_tile(const float *zs, float *pos, const float *ulut, const pl_norm_req &rq,
                                            unsigned short *out, const int tid, const NormArgs *peers = nullptr,
                                            float *rowtab = nullptr)
{
    static_assert(!C4 || (!FAST && !PUSH), "RGBA8 normals: exact arithmetic, no push");
    using namespace plf2;
    using GEO = NGeo<TW>;
    constexpr int W = GEO::W, GWP = GEO::GWP, EPITCH = GEO::EPITCH, UL = GEO::ULUT;
    constexpr int GPAIRS = GEO::GPAIRS, XPAIRS = GEO::XPAIRS, PP = GEO::POS_PLANE;

    const float D = rq.deform[2], R = rq.deform[3];
    const float x0f = rq.deform[0], y0f = rq.deform[1];
    const float s = rq.smooth;
    const float w00 = rq.w2t[0], w01 = rq.w2t[1], w02 = rq.w2t[2];
    const float w10 = rq.w2t[3], w11 = rq.w2t[4], w12 = rq.w2t[5];
    const bool w2t_identity = w00 == 1.0f && w01 == 0.0f && w02 == 0.0f && w10 == 0.0f && w11 == 1.0f && w12 == 0.0f;

    /* Thread -> (column, row lane) of both phases.  A warp's lanes must not straddle two rows inside a
     * half-warp: the two rows' 8-byte accesses would then share banks (2.5 instead of 2 wavefronts per
     * LDS.64 / STS.64: measured, and shared-memory wavefronts are as scarce as fp32 issue here).  So the
     * first 48 columns (three half-warps) of every row lane go to threads 0 .. 48*LANES-1, and the one
     * or two columns that are left to the threads after them. */
    constexpr int CMAIN = 48;
    constexpr int PR = (NT / GPAIRS) < (NT / CMAIN) ? (NT / GPAIRS) : (NT / CMAIN);   /* grid rows per pass */
    constexpr int KR = (NT / XPAIRS) < (NT / CMAIN) ? (NT / XPAIRS) : (NT / CMAIN);   /* block rows per pass */
    static_assert(GPAIRS >= CMAIN && XPAIRS >= CMAIN && PR * GPAIRS <= NT && KR * XPAIRS <= NT, "thread mapping");
    /* position phase: thread = (pair column, row lane) */
    int rl0, pc;
    if (tid < CMAIN * PR) { rl0 = tid / CMAIN; pc = tid - rl0 * CMAIN; }
    else { const int e = tid - CMAIN * PR; rl0 = e / (GPAIRS - CMAIN); pc = CMAIN + e - rl0 * (GPAIRS - CMAIN); }
    const int pc2 = 2 * pc;
    const bool pc_ok = rl0 < PR;
    const float *uA = ulut + pc2, *uB = ulut + UL + pc2;
    const F2 uFA = *reinterpret_cast<const F2 *>(uA), uFB = *reinterpret_cast<const F2 *>(uB);
    /* normal phase: thread = (block column, block row lane) */
    int kl0, xc;
    if (tid < CMAIN * KR) { kl0 = tid / CMAIN; xc = tid - kl0 * CMAIN; }
    else { const int e = tid - CMAIN * KR; kl0 = e / (XPAIRS - CMAIN); xc = CMAIN + e - kl0 * (XPAIRS - CMAIN); }
    const int xc2 = 2 * xc;
    const bool xc_ok = kl0 < KR;

    /* Flat terrain: p = (x0 + D u, y0 + D v, h), so the x difference of two horizontal neighbours depends
     * on the column only, the y difference of two vertical neighbours on the row only, and the other
     * cross terms are exact zeros.  Only h needs a plane; the two difference tables live in the third
     * plane (unused then): dxa[k] = x(X = k - 1) - x(X = k - 3) (texel k - 2), dxb[k] = dxa[k + 1] for
     * aligned pairs at odd texels, dyt likewise with y0.  Same fp32 operations as the general formula. */
    float *dxa = pos + 2 * PP, *dxb = dxa + GWP, *dyt = dxb + GWP;
    if (!SPHERE) {
        for (int k = tid; k < W + 4; k += NT) {
            const bool in = k >= 1 && k + 1 < UL;
            const float hi_u = in ? ulut[k + 1] : 0.0f, lo_u = in ? ulut[k - 1] : 0.0f;
            const float ddx = fmaf(D, hi_u, x0f) - fmaf(D, lo_u, x0f);
            const float ddy = fmaf(D, hi_u, y0f) - fmaf(D, lo_u, y0f);
            dxa[k] = ddx;
            if (k >= 1) dxb[k - 1] = ddx;
            dyt[k] = ddy;
        }
    }

    /* FAST sphere positions: the rows' end values of the 7 bilinear forms (see above).  Corner c of the patch has
     * weight alpha_c = (U V, u V, U v, u v)_c; with Q_c = L_c * (N column c | C column c | 1):
     *   row start A = V Q_0 + v Q_2 (u = 0), row end B = V Q_1 + v Q_3 (u = 1), point = A + u (B - A).
     * Layout of a row: [A_n0 A_n1 A_n2 A_c0][A_c1 A_c2 A_den -][d_n0 d_n1 d_n2 d_c0][d_c1 d_c2 d_den -] */
    const bool fastpos = FAST && SPHERE && s == 1.0f;
    auto row_table = [&](int band) {
        const int y_begin = band * kTileBand;
        const int rows = band == GEO::NBANDS - 1 ? GEO::LAST_ROWS : kTileBand;
        const int r = (band == 0 ? 0 : 2) + tid;
        if (r < rows + 2) {
            const float v = ulut[y_begin + r + 1], V = 1.0f - v;
            float *t = rowtab + r * 16;
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                const float *M = k < 3 ? rq.verticals + 4 * k : rq.corners + 4 * (k - 3);
                const float q0 = k < 6 ? rq.norms[0] * M[0] : rq.norms[0], q1 = k < 6 ? rq.norms[1] * M[1] : rq.norms[1];
                const float q2 = k < 6 ? rq.norms[2] * M[2] : rq.norms[2], q3 = k < 6 ? rq.norms[3] * M[3] : rq.norms[3];
                const float A = fmaf(v, q2, V * q0), B = fmaf(v, q3, V * q1);
                t[k] = A;
                t[8 + k] = B - A;
            }
        }
    };
    if (fastpos) {
        row_table(0);
        __syncthreads();
    }

#pragma unroll 1
    for (int band = 0; band < GEO::NBANDS; ++band) {
        const int y_begin = band * kTileBand;
        const int rows = band == GEO::NBANDS - 1 ? GEO::LAST_ROWS : kTileBand;
        /* grid rows of the band: local r = 0 .. rows+1 (grid row g = y_begin + r is Y = g - 1);
         * rows 0, 1 of every band but the first were computed by the band before */
        const int r_lo = band == 0 ? 0 : 2;
        if (band != 0) {
            __syncthreads();   /* the band before has read its positions */
            constexpr int N4 = 2 * GWP / 4;
            if (tid < (SPHERE ? 3 : 1) * N4) {
                const int pl = tid / N4, c = tid - pl * N4;
                float4 *dst = reinterpret_cast<float4 *>(pos + pl * PP) + c;
                *dst = *reinterpret_cast<const float4 *>(pos + pl * PP + kTileBand * GWP + 4 * c);
            }
            __syncthreads();   /* rows kTileBand, kTileBand+1 are about to be rewritten */
        }
        /* ---- world position of the new grid points of the band, two per thread ----
         * thread = (pair column pc, row lane rl0): its columns never change, rows advance by PR per pass,
         * so every address below is the previous one plus a constant */
        {
            const int r_hi = rows + 2;
            int r = r_lo + rl0;
            const float *zrow = zs + (y_begin + r + 1) * EPITCH + pc2;   /* row Y + 2 of the tile, column pc2 */
            const float *vp = ulut + y_begin + r + 1;                    /* Y = y_begin + r - 1 */
            float *o3 = pos + r * GWP + pc2 + 2;
            const float4 *rt = reinterpret_cast<const float4 *>(rowtab + r * 16);
            if (pc_ok)
            for (; r < r_hi; r += PR, zrow += PR * EPITCH, vp += PR, o3 += PR * GWP, rt += PR * 4) {
                const int sh = row_shift(r);
                /* grid points gx, gx+1 with gx = pc2 - sh (X = gx - 1, gx): elevation texels (gx + 1, Y + 2),
                 * (gx + 2, Y + 2).  Column -1 / GW of a shifted or last pair is a pad: it reads inside the
                 * staged plane (or the guard floats in front of it) and is never used. */
                const float *row1 = zrow - sh, *row0 = row1 - EPITCH;
                F2 h;
                if (!LINEAR) {
                    h = make_float2(row1[1], row1[2]);
                } else {
                    const F2 t00 = make_float2(row0[0], row0[1]);
                    const F2 t10 = make_float2(t00.y, row0[2]);
                    const F2 t01 = make_float2(row1[0], row1[1]);
                    const F2 t11 = make_float2(t01.y, row1[2]);
                    h = fma2(bc(0.5625f), t11, fma2(bc(0.1875f), t01, fma2(bc(0.1875f), t10, mul2(bc(0.0625f), t00))));
                }
                /* u of X = gx - 1, gx: table index X + 2; the pair is aligned in the first copy when gx is odd,
                 * in the second (one entry further) when it is even */
                F2 u = bc(0.0f);
                float v = 0.0f;
                if (SPHERE && FAST && fastpos) {
                    u = sh ? uFA : uFB;            /* the thread's columns never change: both pairs live in registers */
                } else if (SPHERE) {
                    u = *reinterpret_cast<const F2 *>(sh ? uA : uB);
                    v = *vp;
                }
                F2 qx, qy, qz;
                if (!SPHERE) {
                    qx = qy = qz = h;
                } else if (FAST && fastpos) {
                    const float4 A0 = rt[0], A1 = rt[1], D0 = rt[2], D1 = rt[3];
                    const F2 den = fma2(u, bc(D1.z), bc(A1.z));
                    const F2 r0 = make_float2(plfp::rcp_seed(den.x), plfp::rcp_seed(den.y));
                    const F2 rden = fma2(r0, fma2(r0, neg(den), bc(1.0f)), r0);
                    const F2 nx = fma2(u, bc(D0.x), bc(A0.x)), ny = fma2(u, bc(D0.y), bc(A0.y)), nz = fma2(u, bc(D0.z), bc(A0.z));
                    const F2 cx = fma2(u, bc(D0.w), bc(A0.w)), cy = fma2(u, bc(D1.x), bc(A1.x)), cz = fma2(u, bc(D1.y), bc(A1.y));
                    qx = mul2(fma2(h, nx, cx), rden);
                    qy = mul2(fma2(h, ny, cy), rden);
                    qz = mul2(fma2(h, nz, cz), rden);
                } else {
                    const F2 U = sub2(bc(1.0f), u);
                    const float V = 1.0f - v;
                    const F2 a0 = mul2(U, bc(V)), a1 = mul2(u, bc(V)), a2 = mul2(U, bc(v)), a3 = mul2(u, bc(v));
                    const F2 l0 = mul2(a0, bc(rq.norms[0])), l1 = mul2(a1, bc(rq.norms[1]));
                    const F2 l2 = mul2(a2, bc(rq.norms[2])), l3 = mul2(a3, bc(rq.norms[3]));
                    const F2 den = fma2(a3, bc(rq.norms[3]), fma2(a2, bc(rq.norms[2]), fma2(a1, bc(rq.norms[1]), l0)));
                    const F2 rden = rcp_rn2(den);
                    const F2 p0 = div_rn2(l0, den, rden), p1 = div_rn2(l1, den, rden);
                    const F2 p2q = div_rn2(l2, den, rden), p3 = div_rn2(l3, den, rden);
#define ROW4(M, r_) fma2(bc(M[4 * (r_) + 3]), p3, fma2(bc(M[4 * (r_) + 2]), p2q, fma2(bc(M[4 * (r_) + 1]), p1, mul2(bc(M[4 * (r_)]), p0))))
                    const F2 upx = ROW4(rq.verticals, 0), upy = ROW4(rq.verticals, 1), upz = ROW4(rq.verticals, 2);
                    F2 hp = h;
                    if (s != 1.0f) {   /* tile-uniform: levels whose quad is larger than R/64 */
                        const F2 len = sqrt_rn2(plf2::dot3(upx, upy, upz, upx, upy, upz));
                        const F2 kk = fma2(bc(1.0f), bc(s), mul2(len, bc(1.0f - s)));   /* mix(len, 1, s) */
                        hp = div_rn2(fma2(bc(R), sub2(bc(1.0f), kk), h), kk, rcp_rn2(kk));
                    }
                    qx = fma2(hp, upx, ROW4(rq.corners, 0));
                    qy = fma2(hp, upy, ROW4(rq.corners, 1));
                    qz = fma2(hp, upz, ROW4(rq.corners, 2));
#undef ROW4
                }
                /* stored at column gx + shift + 2 (the +2 keeps the odd-start pad at a non-negative, even slot) */
                if (SPHERE) {
                    *reinterpret_cast<F2 *>(o3) = qx;
                    *reinterpret_cast<F2 *>(o3 + PP) = qy;
                    *reinterpret_cast<F2 *>(o3 + 2 * PP) = qz;
                } else {
                    *reinterpret_cast<F2 *>(o3) = qz;     /* flat: the height plane only */
                }
            }
        }
        __syncthreads();
        if (FAST && fastpos && band + 1 < GEO::NBANDS) row_table(band + 1);   /* read after the next band's barriers */

        /* ---- normals of the band: a 2 x 2 block of texels per thread ------------- */
        /* thread = (block column xc, block row lane kl0): columns fixed, block rows advance by KR per pass */
        const int nbr = (rows + 1) >> 1;
        int k = kl0;
        const float *c1 = pos + (2 * k + 1) * GWP + xc2 + 2;       /* grid column x of row g1 (aligned pair) */
        unsigned short *ob = out + (y_begin + 2 * k) * W + xc2;
        if (xc_ok)
        for (; k < nbr; k += KR, c1 += 2 * KR * GWP, ob += 2 * KR * W) {
            const int odd = k & 1;
            const int x = xc2 - odd;                         /* -1, 1, 3, .. on odd block rows, 0, 2, .. on even ones */
            const int ry = 2 * k;
            /* texel (x, ry) is grid column x + 1 of grid row ry + 1.  Centre rows g1 = ry+1, g2 = ry+2 have
             * shift `odd`; outer rows g0 = ry, g3 = ry+3 have shift 1 - odd.  Stored column = grid column + shift + 2,
             * so row g1 starts at x + odd + 2 = xc2 + 2 and row g0 (grid column x + 1) at xc2 + 4 - 2 odd. */
            const float *c0 = c1 - GWP + 2 - 2 * odd;
            F2 d1[3], e1[3], d2[3], e2[3];
            if (!SPHERE) {
                const F2 l1 = *reinterpret_cast<const F2 *>(c1), r1 = *reinterpret_cast<const F2 *>(c1 + 2);
                const F2 l2 = *reinterpret_cast<const F2 *>(c1 + GWP), r2 = *reinterpret_cast<const F2 *>(c1 + GWP + 2);
                const F2 m0 = *reinterpret_cast<const F2 *>(c0), m3 = *reinterpret_cast<const F2 *>(c0 + 3 * GWP);
                const F2 m1 = make_float2(l1.y, r1.x), m2 = make_float2(l2.y, r2.x);
                const F2 ddx = *reinterpret_cast<const F2 *>(odd ? dxb + xc2 : dxa + xc2 + 2);
                const float *dyp = dyt + y_begin + ry + 2;
                d1[0] = d2[0] = ddx;          d1[1] = d2[1] = bc(0.0f);   d1[2] = sub2(r1, l1);  d2[2] = sub2(r2, l2);
                e1[0] = e2[0] = bc(0.0f);     e1[1] = bc(dyp[0]);         e2[1] = bc(dyp[1]);
                e1[2] = sub2(m2, m0);         e2[2] = sub2(m3, m1);
            } else
#pragma unroll
            for (int cpt = 0; cpt < 3; ++cpt) {
                const F2 l1 = *reinterpret_cast<const F2 *>(c1 + cpt * PP);                 /* g1: x, x+1 */
                const F2 r1 = *reinterpret_cast<const F2 *>(c1 + cpt * PP + 2);             /* g1: x+2, x+3 */
                const F2 l2 = *reinterpret_cast<const F2 *>(c1 + cpt * PP + GWP);           /* g2: x, x+1 */
                const F2 r2 = *reinterpret_cast<const F2 *>(c1 + cpt * PP + GWP + 2);       /* g2: x+2, x+3 */
                const F2 m0 = *reinterpret_cast<const F2 *>(c0 + cpt * PP);                 /* g0: x+1, x+2 */
                const F2 m3 = *reinterpret_cast<const F2 *>(c0 + cpt * PP + 3 * GWP);       /* g3: x+1, x+2 */
                const F2 m1 = make_float2(l1.y, r1.x), m2 = make_float2(l2.y, r2.x);        /* g1, g2: x+1, x+2 */
                d1[cpt] = sub2(r1, l1);   /* texel row ry:   right - left */
                d2[cpt] = sub2(r2, l2);   /* texel row ry+1 */
                e1[cpt] = sub2(m2, m0);   /* texel row ry:   up - down */
                e2[cpt] = sub2(m3, m1);   /* texel row ry+1 */
            }
            unsigned int rg[2][2];   /* [row][texel]: r | g << 8 */
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const F2 *d = half ? d2 : d1, *e = half ? e2 : e1;
                F2 nx, ny, nz;
                if (SPHERE) {
                    nx = fma2(d[1], e[2], neg(mul2(d[2], e[1])));
                    ny = fma2(d[2], e[0], neg(mul2(d[0], e[2])));
                    nz = fma2(d[0], e[1], neg(mul2(d[1], e[0])));
                } else {   /* the same cross product with its exact zeros removed (d.y = e.x = 0) */
                    nx = neg(mul2(d[2], e[1]));
                    ny = neg(mul2(d[0], e[2]));
                    nz = mul2(d[0], e[1]);
                }
                if (FAST) {
                    const F2 n2 = plf2::dot3(nx, ny, nz, nx, ny, nz);
                    const F2 i2 = mul2(make_float2(plfp::rsqrt_seed(n2.x), plfp::rsqrt_seed(n2.y)), bc(127.5f));
                    F2 tx = nx, ty = ny;
                    if (SPHERE || !w2t_identity) {
                        tx = plf2::dot3(bc(w00), bc(w01), bc(w02), nx, ny, nz);
                        ty = plf2::dot3(bc(w10), bc(w11), bc(w12), nx, ny, nz);
                    }
                    const F2 r = fma2(tx, i2, bc(127.5f)), g = fma2(ty, i2, bc(127.5f));
                    const unsigned int rx = __float_as_uint(r.x + 12582912.0f), ry = __float_as_uint(r.y + 12582912.0f);
                    const unsigned int gx = __float_as_uint(g.x + 12582912.0f), gy = __float_as_uint(g.y + 12582912.0f);
                    rg[half][0] = __byte_perm(rx, gx, 0x0040u);
                    rg[half][1] = __byte_perm(ry, gy, 0x0040u);
                    continue;
                }
                const F2 inv = rcp_rn2(sqrt_rn2(plf2::dot3(nx, ny, nz, nx, ny, nz)));
                nx = mul2(nx, inv); ny = mul2(ny, inv); nz = mul2(nz, inv);
                /* worldToTangentFrame; the identity of a flat terrain (NormalProducer.cpp:279-280) returns nx, ny
                 * themselves (up to the sign of a zero, which the unorm8 conversion does not see) */
                F2 tx = nx, ty = ny;
                if (SPHERE || !w2t_identity) {
                    tx = plf2::dot3(bc(w00), bc(w01), bc(w02), nx, ny, nz);
                    ty = plf2::dot3(bc(w10), bc(w11), bc(w12), nx, ny, nz);
                }
                /* unorm8: round(clamp(v * 0.5 + 0.5, 0, 1) * 255), NaN -> 0 */
                const F2 r = mul2(make_float2(__saturatef(fmaf(tx.x, 0.5f, 0.5f)), __saturatef(fmaf(tx.y, 0.5f, 0.5f))), bc(255.0f));
                const F2 g = mul2(make_float2(__saturatef(fmaf(ty.x, 0.5f, 0.5f)), __saturatef(fmaf(ty.y, 0.5f, 0.5f))), bc(255.0f));
                /* round to nearest even by adding 1.5 * 2^23 (the sum's ulp is 1; scalar adds: ptxas would contract a
                 * packed multiply + add into one FFMA2, a single rounding): the byte is the low byte of the sum's
                 * bits, one PRMT packs r | g << 8.  No conversion instruction (XU pipe) per texel. */
                const unsigned int rx = __float_as_uint(r.x + 12582912.0f), ry = __float_as_uint(r.y + 12582912.0f);
                const unsigned int gx = __float_as_uint(g.x + 12582912.0f), gy = __float_as_uint(g.y + 12582912.0f);
                rg[half][0] = __byte_perm(rx, gx, 0x0040u);
                rg[half][1] = __byte_perm(ry, gy, 0x0040u);
                if (C4) {
                    /* the coarse normal in the upper two bytes; texels outside the tile are not stored */
                    const int yy = y_begin + 2 * k + half;
                    const int x0c = min(max(x, 0), W - 1), x1c = min(x + 1, W - 1), yc = min(yy, W - 1);
                    rg[half][0] = (rg[half][0] & 0xffffu) | (coarse_normal_ba(*peers, rq, W, SPHERE, x0c, yc, tx.x, ty.x) << 16);
                    rg[half][1] = (rg[half][1] & 0xffffu) | (coarse_normal_ba(*peers, rq, W, SPHERE, x1c, yc, tx.y, ty.y) << 16);
                }
            }
            const bool px0 = x >= 0, px1 = x + 1 < W, py1 = ry + 1 < rows;
            if (C4) {
                unsigned int *o = reinterpret_cast<unsigned int *>(out) + (y_begin + 2 * k) * W + x;
                if (px0) o[0] = rg[0][0];
                if (px1) o[1] = rg[0][1];
                if (px0 && py1) o[W] = rg[1][0];
                if (px1 && py1) o[W + 1] = rg[1][1];
            } else {
                unsigned short *o = ob - odd;
                if (px0) o[0] = (unsigned short) rg[0][0];
                if (px1) o[1] = (unsigned short) rg[0][1];
                if (px0 && py1) o[W] = (unsigned short) rg[1][0];
                if (px1 && py1) o[W + 1] = (unsigned short) rg[1][1];
            }
        }
    }
    if (PUSH) {
        /* The finished tile also goes into every peer's copy of the pool: the gather of finished tiles rides on
         * the kernel that makes them, tile by tile, instead of a collective after it.  The texels are read back
         * from this GPU's L2 (they were stored a moment ago) and leave as 16-byte stores, 512 contiguous bytes
         * per warp instruction: NVLink moves those at full packet size, the 2-byte stores of the producing loop
         * it would not.  The 16-byte tail may run into the slot's padding, never into the next slot. */
        __syncthreads();
        constexpr int N16 = (W * W * 2 + 15) / 16;
        const uint4 *src = reinterpret_cast<const uint4 *>(out);
        for (int i = tid; i < N16; i += NT) {
            const uint4 v = __ldcg(src + i);
            for (int p = 0; p < peers->npeers; ++p)
                reinterpret_cast<uint4 *>(reinterpret_cast<char *>(out) + peers->peer_delta[p])[i] = v;
        }
    }
}

/* ------------------------------------------------------------------------
 * PL_ARITH_FAST, register form (normal_tile_reg): the normal pass of a tile WITHOUT position planes in
 * shared memory and without a CTA barrier.  The banded form above is bound by the shared-memory pipe
 * (profiles/pair_r2b_*: 74 % of its peak -- every position is stored once and loaded 4.5 times) and by the
 * barriers between its phases.  Here a warp owns a strip of texel rows and walks down it; a lane owns FOUR
 * adjacent grid columns (25 lanes cover the 99 grid columns of a tile) and keeps the world positions of the last
 * three grid rows in registers:
 *   - vertical neighbours (up - down) come from the lane's own registers,
 *   - horizontal neighbours are the lane's own columns, except the pair right of its last column, which comes
 *     from the next lane by SHFL (6 per grid row instead of 6 STS.64 + 18 LDS.64 per 2x2 block),
 *   - the row table (the 7 bilinear forms of the FAST position, see normal_tile) is built by the warp for its own
 *     rows (__syncwarp only) from the 28 corner values normal_reg_qtab prepares once per tile; the tangent-frame
 *     rotation is folded into them: W (N a) x W (N b) = W (a x b) for the rotation W = worldToTangentFrame, so
 *     the positions are evaluated IN the tangent frame and the cross product needs no matrix product afterwards,
 *   - flat terrains: only the height is a position; the x difference per column and the y difference per row are
 *     the same fp32 expressions as in the banded form (bit-identical results to it).
 * Heights are fetched exactly as in the banded form (same operations, same order).
 * Used when the tile qualifies (sphere: smoothstep factor exactly 1; flat: identity tangent frame); other tiles of
 * a FAST scene take the banded form.
 * ------------------------------------------------------------------------ */
template <int TW, int NT>
struct RGeo {
    static constexpr int NWARP = NT / 32;
    static constexpr int BASE = TW / NWARP, REM = TW - BASE * NWARP;   /* texel rows per warp: BASE, the first REM warps one more */
    static constexpr int TROWS = BASE + 3;                             /* grid rows of a warp's strip */
    static constexpr int TAB = NWARP * TROWS * 16;                     /* floats: 16 per grid row */
    static constexpr int LANES = (TW + 2 + 3) / 4;                     /* lanes that own grid columns */
    static_assert(LANES <= 32 && 4 * LANES + 1 <= NGeo<TW>::EPITCH, "one warp spans the tile; reads stay inside a staged row");
};

/* The 28 corner values of the 7 bilinear forms, once per tile (any 28 threads; visible after the next barrier):
 * q[k][c] = L_c * (W N)[k][c] (k = 0..2), L_c * (W C)[k-3][c] (k = 3..5), L_c (k = 6) for patch corner c */
__device__ __forceinline__ void normal_reg_qtab(float *qtab, const pl_norm_req &rq, const int t)
{
    if (t < 28) {
        const int k = t >> 2, c = t & 3;
        float m = 1.0f;
        if (k < 6) {
            const float *M = k < 3 ? rq.verticals : rq.corners;
            const float *w = rq.w2t + 3 * (k < 3 ? k : k - 3);
            m = fmaf(w[2], M[8 + c], fmaf(w[1], M[4 + c], w[0] * M[c]));
        }
        qtab[t] = rq.norms[c] * m;
    }
}

template <int NC> struct RegRow { plf2::F2 p[NC][2]; };   /* positions of the lane's 4 grid columns: [component][columns 0,1 | 2,3] */

__device__ __forceinline__ bool normal_reg_ok(const pl_norm_req &rq, const bool sphere)
{
    if (sphere) return rq.smooth == 1.0f;
    return rq.w2t[0] == 1.0f && rq.w2t[1] == 0.0f && rq.w2t[2] == 0.0f && rq.w2t[3] == 0.0f && rq.w2t[4] == 1.0f && rq.w2t[5] == 0.0f;
}

template <int TW, bool SPHERE, bool LINEAR, int NT>
__device__ __forceinline__ void normal_tile_reg(const float *zs, float *tab, const float *qtab, const float *ulut,
                                                const pl_norm_req &rq, unsigned short *out, const int tid)
{
    using namespace plf2;
    using GEO = NGeo<TW>;
    using RG = RGeo<TW, NT>;
    constexpr int W = GEO::W, EP = GEO::EPITCH, NC = SPHERE ? 3 : 1;
    const int lane = tid & 31, wid = tid >> 5;
    const int lc = min(lane, RG::LANES - 1);           /* lanes past the last column group repeat it (their stores are off) */
    const int y0 = wid * RG::BASE + min(wid, RG::REM);  /* first texel row of the warp's strip */
    const int n = RG::BASE + (wid < RG::REM ? 1 : 0);   /* texel rows of the strip; grid rows y0 .. y0 + n + 1 */
    float *wt = tab + wid * (RG::TROWS * 16);

    if (SPHERE) {
        /* row table of the strip: [den A, den D, n0 A, n0 D][c0 A, c0 D, n1 A, n1 D][c1 A, c1 D, n2 A, n2 D][c2 A, c2 D, -, -]
         * (value at u = 0 and difference to u = 1 of each form, in the order a grid point consumes them) */
        if (lane < n + 2) {
            const float v = ulut[y0 + lane + 1], V = 1.0f - v;   /* grid row g = y0 + lane is Y = g - 1 */
            const float4 *q4 = reinterpret_cast<const float4 *>(qtab);
            float *t = wt + lane * 16;
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                const float4 q = q4[k];
                const float A = fmaf(v, q.z, V * q.x), B = fmaf(v, q.w, V * q.y);
                const int at = k == 6 ? 0 : (k < 3 ? 2 + 4 * k : 4 + 4 * (k - 3));
                t[at] = A;
                t[at + 1] = B - A;
            }
        }
        __syncwarp();
    }

    /* the lane's grid columns gx = 4 lc + j (X = gx - 1): u = ulut[gx + 1] */
    const F2 u01 = make_float2(ulut[4 * lc + 1], ulut[4 * lc + 2]), u23 = make_float2(ulut[4 * lc + 3], ulut[4 * lc + 4]);
    /* flat: x difference of the horizontal neighbours of texel x = 4 lc + j (see normal_tile) */
    F2 ddxA = bc(0.0f), ddxB = bc(0.0f);
    const float Dq = rq.deform[2], x0f = rq.deform[0], y0f = rq.deform[1];
    if (!SPHERE) {
        float dd[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int x = 4 * lc + j;
            dd[j] = fmaf(Dq, ulut[min(x + 3, GEO::ULUT - 1)], x0f) - fmaf(Dq, ulut[x + 1], x0f);
        }
        ddxA = make_float2(dd[0], dd[1]);
        ddxB = make_float2(dd[2], dd[3]);
    }

    /* texels of the lane: x = 4 lane + (0, 1 | 2, 3) = centres at grid columns 4 lane + (1, 2 | 3, 4) */
    const bool ok0 = 4 * lane < W, ok1 = 4 * lane + 1 < W, ok2 = 4 * lane + 2 < W, ok3 = 4 * lane + 3 < W;
    const float *zp = zs + y0 * EP + 4 * lc;            /* zm row g, columns gx .. gx + 4 */
    const float *rtp = wt;
    unsigned short *o = out + y0 * W + 4 * lane;

    /* world positions of the lane's four points of the next grid row (rows g, g + 1 of the zm plane; both are
     * loaded here: carrying row g + 1 over to the next step costs more registers than its two loads) */
    auto positions = [&](RegRow<NC> &up) {
        const float4 zn = *reinterpret_cast<const float4 *>(zp + EP);
        const float zn4 = zp[EP + 4];
        F2 h01, h23;
        if (!LINEAR) {
            h01 = make_float2(zn.y, zn.z);
            h23 = make_float2(zn.w, zn4);
        } else {
            const float4 zq = *reinterpret_cast<const float4 *>(zp);
            const float z4 = zp[4];
            h01 = fma2(bc(0.5625f), make_float2(zn.y, zn.z), fma2(bc(0.1875f), make_float2(zn.x, zn.y),
                       fma2(bc(0.1875f), make_float2(zq.y, zq.z), mul2(bc(0.0625f), make_float2(zq.x, zq.y)))));
            h23 = fma2(bc(0.5625f), make_float2(zn.w, zn4), fma2(bc(0.1875f), make_float2(zn.z, zn.w),
                       fma2(bc(0.1875f), make_float2(zq.w, z4), mul2(bc(0.0625f), make_float2(zq.z, zq.w)))));
        }
        zp += EP;
        if (SPHERE) {
            const float4 *rt = reinterpret_cast<const float4 *>(rtp);
            rtp += 16;
            const float4 T0 = rt[0], T1 = rt[1], T2 = rt[2], T3 = rt[3];
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const F2 u = p ? u23 : u01, h = p ? h23 : h01;
                const F2 den = fma2(u, bc(T0.y), bc(T0.x));
                const F2 r0 = make_float2(plfp::rcp_seed(den.x), plfp::rcp_seed(den.y));
                const F2 rden = fma2(r0, fma2(r0, neg(den), bc(1.0f)), r0);
                up.p[0][p] = mul2(fma2(h, fma2(u, bc(T0.w), bc(T0.z)), fma2(u, bc(T1.y), bc(T1.x))), rden);
                up.p[NC > 1 ? 1 : 0][p] = mul2(fma2(h, fma2(u, bc(T1.w), bc(T1.z)), fma2(u, bc(T2.y), bc(T2.x))), rden);
                up.p[NC > 2 ? 2 : 0][p] = mul2(fma2(h, fma2(u, bc(T2.w), bc(T2.z)), fma2(u, bc(T3.y), bc(T3.x))), rden);
            }
        } else {
            up.p[0][0] = h01;
            up.p[0][1] = h23;
        }
    };

    /* normals of one texel row: centre grid row `mid`, the rows below `dn` and above `up`.  Texel pair A (centres at
     * the lane's columns 1, 2) needs the lane's own registers only; pair B (columns 3, 4) takes column 4 -- the next
     * lane's column 0 -- by SHFL: the centre row's pair for the x difference, and the finished y difference */
    auto normals = [&](const int y, const RegRow<NC> &dn, const RegRow<NC> &mid, const RegRow<NC> &up) {
        F2 ddy = bc(0.0f);
        if (!SPHERE) ddy = bc(fmaf(Dq, ulut[y + 3], y0f) - fmaf(Dq, ulut[y + 1], y0f));
        unsigned int rg[4];
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            F2 d[NC], e[NC];
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                if (p == 0) {
                    d[k] = sub2(mid.p[k][1], mid.p[k][0]);
                    e[k] = make_float2(up.p[k][0].y - dn.p[k][0].y, up.p[k][1].x - dn.p[k][1].x);
                } else {
                    const F2 r = make_float2(__shfl_down_sync(0xffffffffu, mid.p[k][0].x, 1), __shfl_down_sync(0xffffffffu, mid.p[k][0].y, 1));
                    const float e4 = __shfl_down_sync(0xffffffffu, up.p[k][0].x - dn.p[k][0].x, 1);
                    d[k] = sub2(r, mid.p[k][1]);
                    e[k] = make_float2(up.p[k][1].y - dn.p[k][1].y, e4);
                }
            }
            F2 nx, ny, nz;
            if (SPHERE) {
                nx = fma2(d[NC > 1 ? 1 : 0], e[NC > 2 ? 2 : 0], neg(mul2(d[NC > 2 ? 2 : 0], e[NC > 1 ? 1 : 0])));
                ny = fma2(d[NC > 2 ? 2 : 0], e[0], neg(mul2(d[0], e[NC > 2 ? 2 : 0])));
                nz = fma2(d[0], e[NC > 1 ? 1 : 0], neg(mul2(d[NC > 1 ? 1 : 0], e[0])));
            } else {   /* d = (ddx, 0, dz), e = (0, ddy, ez): the cross product with its exact zeros removed */
                const F2 ddx = p ? ddxB : ddxA;
                nx = neg(mul2(d[0], ddy));
                ny = neg(mul2(ddx, e[0]));
                nz = mul2(ddx, ddy);
            }
            const F2 n2 = plf2::dot3(nx, ny, nz, nx, ny, nz);
            const F2 i2 = mul2(make_float2(plfp::rsqrt_seed(n2.x), plfp::rsqrt_seed(n2.y)), bc(127.5f));
            /* round to nearest by the 1.5 * 2^23 magic add (the sum's ulp is 1), byte = low byte of the sum's bits */
            const F2 r = add2(fma2(nx, i2, bc(127.5f)), bc(12582912.0f)), g = add2(fma2(ny, i2, bc(127.5f)), bc(12582912.0f));
            rg[2 * p] = __byte_perm(__float_as_uint(r.x), __float_as_uint(g.x), 0x0040u);
            rg[2 * p + 1] = __byte_perm(__float_as_uint(r.y), __float_as_uint(g.y), 0x0040u);
        }
        if (ok0) o[0] = (unsigned short) rg[0];
        if (ok1) o[1] = (unsigned short) rg[1];
        if (ok2) o[2] = (unsigned short) rg[2];
        if (ok3) o[3] = (unsigned short) rg[3];
        o += W;
    };

    RegRow<NC> S0, S1, S2;
    positions(S0);
    positions(S1);
    int y = y0;
    const int y_end = y0 + n;
#pragma unroll 1
    for (; y + 3 <= y_end; y += 3) {
        positions(S2); normals(y, S0, S1, S2);
        positions(S0); normals(y + 1, S1, S2, S0);
        positions(S1); normals(y + 2, S2, S0, S1);
    }
    if (y < y_end) {
        positions(S2); normals(y, S0, S1, S2);
        if (y + 1 < y_end) { positions(S0); normals(y + 1, S1, S2, S0); }
    }
}

/* uv / (tileSDF.x - 1.0) for X = -2 .. (X = -2 and W+1 are pads of odd-start pairs), twice: the second
 * copy starts one entry further, so that a pair starting at an odd index is an aligned pair there */
template <int TW, int NT>
__device__ __forceinline__ void normal_uv_tables(float *ulut, const int tid)
{
    using GEO = NGeo<TW>;
    const float wm1 = (float) GEO::W - 1.0f;
    const float rw = plfp::rcp_rn(wm1);
    for (int q = tid; q < GEO::ULUT; q += NT) {
        const float uq = plfp::div_rn((float) (q - 2), wm1, rw);
        ulut[q] = uq;
        if (q >= 1) ulut[GEO::ULUT + q - 1] = uq;
    }
    if (tid == 0) ulut[2 * GEO::ULUT - 1] = 0.0f;
}

}  // namespace plnorm
#endif
