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
#include "pl_normal_tile.cuh"

namespace {

using namespace plnorm;

constexpr int kThreads = 224;      /* generic kernel */
constexpr int kBandRows = 24;
constexpr int kTileThreads = 256;  /* specialised kernel: one CTA per tile */

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
 * Specialised kernel: compile-time geometry, one CTA per tile; the per-tile
 * device code lives in pl_normal_tile.cuh (shared with pl_pair.cu).
 * ------------------------------------------------------------------------ */
template <int TW, bool SPHERE, bool LINEAR, bool FAST>
__global__ void __launch_bounds__(kTileThreads, 3) normal_kernel_fast(const NormArgs a)
{
    using GEO = NGeo<TW>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *zs = reinterpret_cast<float *>(smem_raw) + 4;             /* the tile's zm plane: EW x EPITCH, behind 4 guard floats */
    float *pos = zs + GEO::EPLANE;                                   /* 3 planes of POS_ROWS x GWP */
    float *ulut = pos + 3 * GEO::POS_PLANE;                          /* 2 x ULUT */
    float *rowtab = ulut + 2 * GEO::ULUT;                            /* PL_ARITH_FAST: row_table of a band */
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
    normal_uv_tables<TW, kTileThreads>(ulut, tid);
    if (FAST && SPHERE) normal_reg_qtab(rowtab, rq, tid);
    __syncthreads();
    mbar_wait(&bar, 0);

    unsigned short *out = reinterpret_cast<unsigned short *>(a.norm + (size_t) rq.out_slot * a.norm_slot_bytes);
    if (FAST && normal_reg_ok(rq, SPHERE)) {
        normal_tile_reg<TW, SPHERE, LINEAR, kTileThreads>(zs, pos, rowtab, ulut, rq, out, tid);
        return;
    }
    normal_tile<TW, SPHERE, LINEAR, kTileThreads, false, FAST>(zs, pos, ulut, rq, out, tid, nullptr, rowtab);
}

}  // namespace


int pl_norm_fill_args(pl_ctx *ctx, const pl_norm_scene *sc, pl_pool *norm, pl_pool *elev, const pl_norm_req *dev_reqs,
                      plnorm::NormArgs &a)
{
    (void) ctx;
    if (norm->kind != PL_POOL_NORM_UN8x2 && norm->kind != PL_POOL_NORM_UN8x4)
        return pl_set_error(PL_ERR_ARG, "normal pool must be RG8 or RGBA8");
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
    a.fast = sc->arith == PL_ARITH_FAST;
    a.norm_slot_bytes = (long long) norm->slot_bytes;
    a.npeers = norm->push == 1 ? norm->npeers : 0;
    for (int p = 0; p < a.npeers; ++p) a.peer_delta[p] = (long long) (norm->peer_base[p] - norm->base);
    if (norm->push == 2) {
        /* multicast: ONE store through the multicast mapping reaches the slot on every GPU of the group (this one
         * included), replicated by the NVLink switch instead of one unicast store per peer */
        a.npeers = 1;
        a.peer_delta[0] = (long long) (norm->mc_base - norm->base);
    }
    return PL_OK;
}

int pl_launch_normal(pl_ctx *ctx, const pl_norm_scene *sc, pl_pool *norm, pl_pool *elev, int n,
                     const pl_norm_req *dev_reqs)
{
    PL_CUDA(cudaSetDevice(ctx->device));
    NormArgs a;
    int rc = pl_norm_fill_args(ctx, sc, norm, elev, dev_reqs, a);
    if (rc) return rc;
    const int GW = a.W + 2;
    const size_t smem = (size_t) (a.max_rows + 3) * a.epitch * 4 + (size_t) 3 * (a.max_rows + 2) * GW * 4 +
                        (size_t) ((GW + 3) & ~3) * 4 + (size_t) ((a.max_rows * a.W * a.channels + 15) & ~15);
    if (a.W == 97 && a.border == 2 && a.channels == 2 && !ctx->force_generic) {
        /* the geometry of every shipped archive: compile-time specialisation */
        using GEO = NGeo<97>;
        const size_t fsmem = GEO::SMEM;
        void (*kern)(NormArgs) =
            a.fast ? (a.sphere ? (a.linear ? normal_kernel_fast<97, true, true, true> : normal_kernel_fast<97, true, false, true>)
                               : (a.linear ? normal_kernel_fast<97, false, true, true> : normal_kernel_fast<97, false, false, true>))
                   : (a.sphere ? (a.linear ? normal_kernel_fast<97, true, true, false> : normal_kernel_fast<97, true, false, false>)
                               : (a.linear ? normal_kernel_fast<97, false, true, false> : normal_kernel_fast<97, false, false, false>));
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
