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

namespace {

constexpr int kThreads = 256;
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
    int channels;            /* 2 (RG8) */
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

__global__ void __launch_bounds__(kThreads) normal_kernel(const NormArgs a)
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
        reinterpret_cast<uchar2 *>(outb)[k] = make_uchar2((unsigned char) r8, (unsigned char) g8);
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

}  // namespace

int pl_launch_normal(pl_ctx *ctx, const pl_norm_scene *sc, pl_pool *norm, pl_pool *elev, int n,
                     const pl_norm_req *dev_reqs)
{
    PL_CUDA(cudaSetDevice(ctx->device));
    if (norm->kind != PL_POOL_NORM_UN8x2)
        return pl_set_error(PL_ERR_ARG, "only RG8 normal pools are implemented (every shipped archive uses RG8)");
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
    a.channels = 2;
    a.nbands = a.W / kBandRows > 0 ? a.W / kBandRows : 1;
    a.max_rows = a.W - (a.nbands - 1) * kBandRows;
    a.norm_slot_bytes = (long long) norm->slot_bytes;
    const int GW = a.W + 2;
    const size_t smem = (size_t) (a.max_rows + 3) * a.epitch * 4 + (size_t) 3 * (a.max_rows + 2) * GW * 4 +
                        (size_t) ((GW + 3) & ~3) * 4 + (size_t) ((a.max_rows * a.W * a.channels + 15) & ~15);
    if (smem > 227 * 1024) return pl_set_error(PL_ERR_ARG, "normal tile_w %d needs %zu bytes of shared memory", a.W, smem);
    /* static + dynamic must stay under the default 48 KB unless opted in */
    if (smem > 40 * 1024)
        PL_CUDA(cudaFuncSetAttribute(normal_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    pl_timing_begin(ctx, PL_K_NORMAL, n);
    normal_kernel<<<n * a.nbands, kThreads, smem, ctx->stream>>>(a);
    pl_timing_end(ctx);
    PL_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return PL_OK;
}
