/*
 * pl_normal.cu -- the batched NormalProducer pass for sm_100a.
 *
 * What the reference draws as one 97x97 quad with normalShader
 * (src/demo/shaders/elevation/normalShader.glsl:60-125; example variants in
 * src/terrain/examples/terrain{1,2}/normalShader.glsl) after
 * NormalProducer::doCreateTile (NormalProducer.cpp:164-289) has set the
 * uniforms.
 *
 * Work split: a CTA owns a band of 24 (last band: the remainder) output rows of
 * one tile; grid = tiles x bands.
 *   1. the zm rows the band needs are contiguous in the pitch-padded plane: ONE
 *      1-D bulk copy (cp.async.bulk, mbarrier-signalled) stages them in shared
 *      memory
 *   2. the shader evaluates getWorldPosition at the four neighbours of every
 *      texel; the position is a pure function of the grid point, so it is
 *      evaluated ONCE per grid point of the band (+1 ring) into shared memory
 *      instead of 4x per texel (this is where the sphere maths lives)
 *   3. each texel takes its four neighbours' positions, forms the normal,
 *      rotates it to the tangent frame and quantises to unorm8 into a shared
 *      staging band
 *   4. the band leaves as ONE bulk store (dense rows: the band is contiguous in
 *      HBM; bands start on 16-byte boundaries because they start on rows that
 *      are multiples of 8)
 *
 * Sampler semantics: the shader fetches elevation at texel
 * centre + 0.25; NEAREST storage -> that texel, LINEAR storage -> the
 * (.25,.75)x(.25,.75) blend of the 2x2 block ending at it.
 *
 * Arithmetic: canonical fp32 order of oracle/orc_fp.h (compiled with
 * --fmad=false, fused operations spelled fmaf) -> bit-identical to the oracle.
 */
#include "pl_internal.h"
#include "pl_fpexact.cuh"
#include "pl_f2.cuh"

namespace {

constexpr int kThreads = 224;
constexpr int kBandRows = 24;

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
    long long norm_slot_bytes;
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
 * Generic kernel: runtime geometry (any tile width / border).
 * ------------------------------------------------------------------------ */
__global__ void __launch_bounds__(kThreads) normal_kernel_generic(const NormArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ pl_norm_req rq;

    const int tid = threadIdx.x;
    const int tile = blockIdx.x / a.nbands, band = blockIdx.x - tile * a.nbands;
    const int W = a.W, b = a.border;
    const int y_begin = band * kBandRows;
    const int rows = band == a.nbands - 1 ? W - y_begin : kBandRows;
    const int GW = W + 2;                /* grid points X = -1 .. W */
    const int grows = rows + 2;          /* grid rows Y = y_begin-1 .. y_begin+rows */

    /* shared memory carve-up (sizes for the largest band) */
    const int zrows_max = a.max_rows + 3;
    float *zs = reinterpret_cast<float *>(smem_raw);                 /* zrows x epitch (16-byte multiple) */
    uint8_t *outb = reinterpret_cast<uint8_t *>(zs + zrows_max * a.epitch);   /* band staging, 16-byte aligned */
    float *px = reinterpret_cast<float *>(outb + ((a.max_rows * W * a.channels + 15) & ~15));   /* grows x GW, x3 */
    float *py = px + (a.max_rows + 2) * GW;
    float *pz = py + (a.max_rows + 2) * GW;
    float *ulut = pz + (a.max_rows + 2) * GW;                        /* GW */

    {   /* request -> shared (240 bytes) */
        const int *src = reinterpret_cast<const int *>(a.reqs + tile);
        int *dst = reinterpret_cast<int *>(&rq);
        if (tid < (int) (sizeof(pl_norm_req) / 4)) dst[tid] = __ldg(src + tid);
    }
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    /* elevation rows needed: texel rows (Y + b) and, for LINEAR, (Y + b - 1) */
    const int zr0 = max(y_begin - 1 + b - 1, 0);
    const int zr1 = min(y_begin + rows + b, a.EW - 1);
    const int zrows = zr1 - zr0 + 1;
    if (tid == 0) {
        const float *src = a.elev + (size_t) rq.elev_slot * 3 * a.eplane + 2 * (size_t) a.eplane + (size_t) zr0 * a.epitch;
        const uint32_t bytes = (uint32_t) (zrows * a.epitch * sizeof(float));
        mbar_expect_tx(&bar, bytes);
        bulk_load(zs, src, bytes, &bar);
    }
    /* uv table: uv / (tileSDF.x - 1.0) for X = -1 .. W */
    for (int k = tid; k < GW; k += kThreads) ulut[k] = (float) (k - 1) / ((float) W - 1.0f);
    __syncthreads();
    mbar_wait(&bar, 0);

    const float D = rq.deform[2], R = rq.deform[3];
    const float x0f = rq.deform[0], y0f = rq.deform[1];
    const bool sphere = R != 0.0f;
    const float s = rq.smooth;

    /* ---- 2. world position of every grid point of the band ---------------- */
    for (int k = tid; k < grows * GW; k += kThreads) {
        const int gy = k / GW, gx = k - gy * GW;
        const int X = gx - 1, Y = y_begin - 1 + gy;
        /* elevation texel (X + b, Y + b) through the storage's filter */
        const int ix = X + b, iy = Y + b;
        float h;
        if (!a.linear) {
            h = zs[(iy - zr0) * a.epitch + ix];
        } else {
            const int ix0 = max(ix - 1, 0), iy0 = max(iy - 1, 0);
            const float t00 = zs[(iy0 - zr0) * a.epitch + ix0], t10 = zs[(iy0 - zr0) * a.epitch + ix];
            const float t01 = zs[(iy - zr0) * a.epitch + ix0], t11 = zs[(iy - zr0) * a.epitch + ix];
            const float fa = 0.75f, fb = 0.75f;
            h = fmaf(fa * fb, t11, fmaf((1.0f - fa) * fb, t01, fmaf(fa * (1.0f - fb), t10, ((1.0f - fa) * (1.0f - fb)) * t00)));
        }
        const float u = ulut[gx], v = ulut[Y + 1];
        float qx, qy, qz;
        if (!sphere) {
            qx = fmaf(D, u, x0f);
            qy = fmaf(D, v, y0f);
            qz = h;
        } else {
            const float U = 1.0f - u, V = 1.0f - v;
            const float a0 = U * V, a1 = u * V, a2 = U * v, a3 = u * v;
            const float den = fmaf(a3, rq.norms[3], fmaf(a2, rq.norms[2], fmaf(a1, rq.norms[1], a0 * rq.norms[0])));
            const float p0 = a0 * rq.norms[0] / den, p1 = a1 * rq.norms[1] / den;
            const float p2 = a2 * rq.norms[2] / den, p3 = a3 * rq.norms[3] / den;
            const float upx = dot4(rq.verticals + 0, p0, p1, p2, p3);
            const float upy = dot4(rq.verticals + 4, p0, p1, p2, p3);
            const float upz = dot4(rq.verticals + 8, p0, p1, p2, p3);
            float hp = h;
            if (s != 1.0f) {
                const float len = sqrtf(dot3(upx, upy, upz, upx, upy, upz));
                const float kk = fmaf(1.0f, s, len * (1.0f - s));   /* mix(len, 1, s) */
                hp = fmaf(R, 1.0f - kk, h) / kk;
            }
            qx = fmaf(hp, upx, dot4(rq.corners + 0, p0, p1, p2, p3));
            qy = fmaf(hp, upy, dot4(rq.corners + 4, p0, p1, p2, p3));
            qz = fmaf(hp, upz, dot4(rq.corners + 8, p0, p1, p2, p3));
        }
        px[k] = qx;
        py[k] = qy;
        pz[k] = qz;
    }
    __syncthreads();

    /* ---- 3. normals of the band ------------------------------------------- */
    const int C = a.channels;
    for (int k = tid; k < rows * W; k += kThreads) {
        const int ry = k / W, x = k - ry * W;
        const int c = (ry + 1) * GW + (x + 1);          /* grid index of (x, y) */
        const float ax = px[c + 1] - px[c - 1], ay = py[c + 1] - py[c - 1], az = pz[c + 1] - pz[c - 1];
        const float cx = px[c + GW] - px[c - GW], cy = py[c + GW] - py[c - GW], cz = pz[c + GW] - pz[c - GW];
        float nx = fmaf(ay, cz, -(az * cy));
        float ny = fmaf(az, cx, -(ax * cz));
        float nz = fmaf(ax, cy, -(ay * cx));
        const float inv = 1.0f / sqrtf(dot3(nx, ny, nz, nx, ny, nz));
        nx *= inv; ny *= inv; nz *= inv;
        const float tx = dot3(rq.w2t[0], rq.w2t[1], rq.w2t[2], nx, ny, nz);
        const float ty = dot3(rq.w2t[3], rq.w2t[4], rq.w2t[5], nx, ny, nz);
        const unsigned int r8 = unorm8(fmaf(tx, 0.5f, 0.5f)), g8 = unorm8(fmaf(ty, 0.5f, 0.5f));
        if (C == 2) {
            reinterpret_cast<uchar2 *>(outb)[k] = make_uchar2((unsigned char) r8, (unsigned char) g8);
        } else {
            /* RGBA8 (tileSDF.z = 1): .zw = the parent's coarse normal, normalShader.glsl:100-114 */
            float ncx = tx, ncy = ty;
            if (rq.parent_slot >= 0) {
                const uchar4 *parent = reinterpret_cast<const uchar4 *>(a.norm + (size_t) rq.parent_slot * a.norm_slot_bytes);
                const int g = a.grid, y = y_begin + ry;
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
            const unsigned int b8 = unorm8(fmaf(ncx, 0.5f, 0.5f)), a8 = unorm8(fmaf(ncy, 0.5f, 0.5f));
            reinterpret_cast<uchar4 *>(outb)[k] = make_uchar4((unsigned char) r8, (unsigned char) g8, (unsigned char) b8, (unsigned char) a8);
        }
    }
    /* generic-proxy writes -> visible to the bulk (async proxy) store */
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    /* ---- 4. one bulk store of the band ------------------------------------- */
    if (tid == 0) {
        uint8_t *dst = a.norm + (size_t) rq.out_slot * a.norm_slot_bytes + (size_t) y_begin * W * C;
        const uint32_t bytes = (uint32_t) ((rows * W * C + 15) & ~15);
        bulk_store(dst, outb, bytes);
    }
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
constexpr int kTileThreads = 256;
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
    static constexpr size_t SMEM = 16 + (size_t) EPLANE * 4 + (size_t) 3 * POS_PLANE * 4 + (size_t) 2 * ULUT * 4;   /* 16: guard floats in front of the zm plane */
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
template <int TW, bool SPHERE, bool LINEAR, int NT>
__device__ __forceinline__ void normal_tile(const float *zs, float *pos, const float *ulut, const pl_norm_req &rq,
                                            unsigned short *out, const int tid)
{
    using namespace plf2;
    using GEO = NGeo<TW>;
    constexpr int W = GEO::W, GWP = GEO::GWP, EPITCH = GEO::EPITCH, UL = GEO::ULUT;
    constexpr int GPAIRS = GEO::GPAIRS, XPAIRS = GEO::XPAIRS, PP = GEO::POS_PLANE;

    const float D = rq.deform[2], R = rq.deform[3];
    const float x0f = rq.deform[0], y0f = rq.deform[1];
    const float s = rq.smooth;
    const float w00 = rq.w2t[0], w01 = rq.w2t[1], w02 = rq.w2t[2];
    const float w10 = rq.w2t[3], w11 = rq.w2t[4], w12 = rq.w2t[5];

    /* position phase: thread = (pair column, row lane) */
    constexpr int PR = NT / GPAIRS;                    /* grid rows per pass */
    const int rl0 = tid / GPAIRS, pc2 = 2 * (tid - rl0 * GPAIRS);
    const bool pc_ok = rl0 < PR;
    const float *uA = ulut + pc2, *uB = ulut + UL + pc2;
    /* normal phase: thread = (block column, block row lane) */
    constexpr int KR = NT / XPAIRS;                    /* block rows per pass */
    const int kl0 = tid / XPAIRS, xc2 = 2 * (tid - kl0 * XPAIRS);
    const bool xc_ok = kl0 < KR;

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
            if (tid < 3 * N4) {
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
            if (pc_ok)
            for (; r < r_hi; r += PR, zrow += PR * EPITCH, vp += PR, o3 += PR * GWP) {
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
                const F2 u = *reinterpret_cast<const F2 *>(sh ? uA : uB);
                const float v = *vp;
                F2 qx, qy, qz;
                if (!SPHERE) {
                    qx = fma2(bc(D), u, bc(x0f));
                    qy = bc(fmaf(D, v, y0f));
                    qz = h;
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
                *reinterpret_cast<F2 *>(o3) = qx;
                *reinterpret_cast<F2 *>(o3 + PP) = qy;
                *reinterpret_cast<F2 *>(o3 + 2 * PP) = qz;
            }
        }
        __syncthreads();

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
                F2 nx = fma2(d[1], e[2], neg(mul2(d[2], e[1])));
                F2 ny = fma2(d[2], e[0], neg(mul2(d[0], e[2])));
                F2 nz = fma2(d[0], e[1], neg(mul2(d[1], e[0])));
                const F2 inv = rcp_rn2(sqrt_rn2(plf2::dot3(nx, ny, nz, nx, ny, nz)));
                nx = mul2(nx, inv); ny = mul2(ny, inv); nz = mul2(nz, inv);
                const F2 tx = plf2::dot3(bc(w00), bc(w01), bc(w02), nx, ny, nz);
                const F2 ty = plf2::dot3(bc(w10), bc(w11), bc(w12), nx, ny, nz);
                /* unorm8: round(clamp(v * 0.5 + 0.5, 0, 1) * 255), NaN -> 0 */
                const F2 r = mul2(make_float2(__saturatef(fmaf(tx.x, 0.5f, 0.5f)), __saturatef(fmaf(tx.y, 0.5f, 0.5f))), bc(255.0f));
                const F2 g = mul2(make_float2(__saturatef(fmaf(ty.x, 0.5f, 0.5f)), __saturatef(fmaf(ty.y, 0.5f, 0.5f))), bc(255.0f));
                rg[half][0] = (unsigned int) __float2int_rn(r.x) + ((unsigned int) __float2int_rn(g.x) << 8);
                rg[half][1] = (unsigned int) __float2int_rn(r.y) + ((unsigned int) __float2int_rn(g.y) << 8);
            }
            unsigned short *o = ob - odd;
            const bool px0 = x >= 0, px1 = x + 1 < W, py1 = ry + 1 < rows;
            if (px0) o[0] = (unsigned short) rg[0][0];
            if (px1) o[1] = (unsigned short) rg[0][1];
            if (px0 && py1) o[W] = (unsigned short) rg[1][0];
            if (px1 && py1) o[W + 1] = (unsigned short) rg[1][1];
        }
    }
}

template <int TW, bool SPHERE, bool LINEAR>
__global__ void __launch_bounds__(kTileThreads, 3) normal_kernel_fast(const NormArgs a)
{
    using GEO = NGeo<TW>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *zs = reinterpret_cast<float *>(smem_raw) + 4;             /* the tile's zm plane: EW x EPITCH, behind 4 guard floats */
    float *pos = zs + GEO::EPLANE;                                   /* 3 planes of POS_ROWS x GWP */
    float *ulut = pos + 3 * GEO::POS_PLANE;                          /* 2 x ULUT */
    __shared__ uint64_t bar;
    __shared__ pl_norm_req rq;

    const int tid = threadIdx.x;
    const int tile = blockIdx.x;
    {
        const int *src = reinterpret_cast<const int *>(a.reqs + tile);
        int *dst = reinterpret_cast<int *>(&rq);
        if (tid < (int) (sizeof(pl_norm_req) / 4)) dst[tid] = __ldg(src + tid);
    }
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        const float *src = a.elev + (size_t) rq.elev_slot * 3 * GEO::EPLANE + 2 * GEO::EPLANE;
        constexpr uint32_t bytes = (uint32_t) (GEO::EPLANE * sizeof(float));
        mbar_expect_tx(&bar, bytes);
        bulk_load(zs, src, bytes, &bar);
    }
    {   /* uv / (tileSDF.x - 1.0) for X = -2 .. (X = -2 and W+1 are pads of odd-start pairs) */
        const float wm1 = (float) GEO::W - 1.0f;
        const float rw = plfp::rcp_rn(wm1);
        for (int q = tid; q < GEO::ULUT; q += kTileThreads) {
            const float uq = plfp::div_rn((float) (q - 2), wm1, rw);
            ulut[q] = uq;
            if (q >= 1) ulut[GEO::ULUT + q - 1] = uq;
        }
        if (tid == 0) ulut[2 * GEO::ULUT - 1] = 0.0f;
    }
    __syncthreads();
    mbar_wait(&bar, 0);

    unsigned short *out = reinterpret_cast<unsigned short *>(a.norm + (size_t) rq.out_slot * a.norm_slot_bytes);
    normal_tile<TW, SPHERE, LINEAR, kTileThreads>(zs, pos, ulut, rq, out, tid);
}

}  // namespace


int pl_launch_normal(pl_ctx *ctx, const pl_norm_scene *sc, pl_pool *norm, pl_pool *elev, int n,
                     const pl_norm_req *dev_reqs)
{
    PL_CUDA(cudaSetDevice(ctx->device));
    if (norm->kind != PL_POOL_NORM_UN8x2 && norm->kind != PL_POOL_NORM_UN8x4)
        return pl_set_error(PL_ERR_ARG, "normal pool must be RG8 or RGBA8");
    NormArgs a;
    a.elev = reinterpret_cast<const float *>(elev->base);
    a.norm = norm->base;
    a.reqs = dev_reqs;
    a.W = norm->tile_w;
    a.EW = elev->tile_w;
    a.epitch = elev->pitch;
    a.eplane = (int) elev->plane_elems;
    a.border = sc->elev_border;
    a.linear = sc->elev_filter == PL_FILTER_LINEAR;
    a.sphere = sc->sphere;
    a.channels = norm->kind == PL_POOL_NORM_UN8x4 ? 4 : 2;
    a.grid = sc->grid;
    a.parent_linear = sc->parent_filter == PL_FILTER_LINEAR;
    a.nbands = a.W / kBandRows > 0 ? a.W / kBandRows : 1;
    a.max_rows = a.W - (a.nbands - 1) * kBandRows;
    a.norm_slot_bytes = (long long) norm->slot_bytes;
    const int GW = a.W + 2;
    const size_t smem = (size_t) (a.max_rows + 3) * a.epitch * 4 + (size_t) 3 * (a.max_rows + 2) * GW * 4 +
                        (size_t) ((GW + 3) & ~3) * 4 + (size_t) ((a.max_rows * a.W * a.channels + 15) & ~15);
    if (a.W == 97 && a.border == 2 && a.channels == 2 && !ctx->force_generic) {
        /* the geometry of every shipped archive: compile-time specialisation */
        using GEO = NGeo<97>;
        const size_t fsmem = GEO::SMEM;
        void (*kern)(NormArgs) = a.sphere ? (a.linear ? normal_kernel_fast<97, true, true> : normal_kernel_fast<97, true, false>)
                                          : (a.linear ? normal_kernel_fast<97, false, true> : normal_kernel_fast<97, false, false>);
        PL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) fsmem));
        pl_timing_begin(ctx, PL_K_NORMAL, n);
        kern<<<n, kTileThreads, fsmem, ctx->stream>>>(a);
        pl_timing_end(ctx);
    } else {
        if (smem > 227 * 1024) return pl_set_error(PL_ERR_ARG, "normal tile_w %d needs %zu bytes of shared memory", a.W, smem);
        /* static + dynamic must stay under the default 48 KB unless opted in */
        if (smem > 40 * 1024)
            PL_CUDA(cudaFuncSetAttribute(normal_kernel_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        pl_timing_begin(ctx, PL_K_NORMAL, n);
        normal_kernel_generic<<<n * a.nbands, kThreads, smem, ctx->stream>>>(a);
        pl_timing_end(ctx);
    }
    PL_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return PL_OK;
}
