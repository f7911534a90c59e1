/*
 * pl_ortho.cu -- OrthoProducer on the device.
 *
 *   terrain/sources/proland/ortho/OrthoProducer.cpp:48-118    createOrthoNoise   -> pl_ortho_noise_init
 *   terrain/sources/proland/ortho/OrthoProducer.cpp:268-372   doCreateTile       -> pl_ortho_make_req, pl_ortho_batch
 *   demo/shaders/ortho/upsampleOrthoShader.glsl:123-158       main()             -> ortho_kernel
 *
 * One CTA per tile.  The parent quadrant ((W/2 + 2)^2 RGBA8 texels, 40 KB for W = 196) is staged in
 * shared memory by one 1-D bulk copy per row (cp.async.bulk, mbarrier-signalled).  A thread produces
 * four horizontally adjacent texels per step: the (9,3,3,1)/16 upsample runs as exact integer
 * arithmetic on two 16-bit lanes per register (the shader's float sums are sums of small integers, so
 * floor(s / 16) == s >> 4), the residual and the noise come in as one 16-byte load each (the noise
 * layers are stored pre-rotated, so every rotation reads rows), and the result leaves as one 16-byte
 * store.  The colour maths follows the oracle's canonical fp32 order (oracle/orc_fp.h) operation by
 * operation; divisions are the correctly rounded ones (pl_fpexact.cuh, constants by the Markstein
 * sequence with the correctly rounded reciprocal).
 *
 * HBM traffic per tile (W = 196): write 153 664 B, parent quadrant read 40 000 B (+ 153 664 B of
 * residual when present); the 24 rotated noise layers (3.7 MB) stay in L2.
 */
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

#include "pl_internal.h"
#include "pl_reqmath.cuh"
#include "pl_fpexact.cuh"

PerlinView pl_host_perlin();   /* pl_hostmath.cu */

static_assert(sizeof(pl_ortho_req) == 64 && sizeof(pl_ortho_scene) == 192, "the layouts proland_b200.py mirrors");

/* ------------------------------------------------------------------ host: noise */

namespace {

struct OrthoLcg {
    uint32_t s;
    explicit OrthoLcg(uint32_t seed) : s(seed) {}
    /* int(frandom(&seed) * 255.0f): 0..254 */
    uint8_t byte()
    {
        s = (uint32_t) (((uint64_t) s * 1103515245ull + 12345ull) & 0x7FFFFFFFull);
        const float u = (float) (s >> 7) / 16777216.0f;
        return (uint8_t) (int) (u * 255.0f);
    }
};

/* one of the four 2-texel border strips of a layer: `outer` takes two values starting at o0 and moving by
 * ostep, `inner` runs over [4, W-4); the mirror image sits at (osum - outer, W-1-inner) */
struct Strip { int bit; bool outer_is_row; int o0, ostep, osum; };

}  // namespace

void pl_host_ortho_noise(int W, uint8_t *out6)
{
    static const int kPattern[6] = { 0, 1, 3, 5, 7, 15 };
    const Strip strips[4] = {
        { 1, true, 2, +1, 3 },                  /* bottom: rows 2,3 mirrored onto rows 1,0          */
        { 2, false, W - 3, -1, 2 * W - 5 },     /* right: columns W-3,W-4 onto W-2,W-1              */
        { 4, true, W - 2, +1, 2 * W - 5 },      /* top: rows W-2,W-1 onto W-3,W-4                   */
        { 8, false, 1, -1, 3 },                 /* left: columns 1,0 onto 2,3                       */
    };
    OrthoLcg interior(1234567u);
    for (int layer = 0; layer < 6; ++layer) {
        uint8_t *n = out6 + (size_t) layer * W * W * 4;
        memset(n, 128, (size_t) W * W * 4);
        for (const Strip &st : strips) {
            OrthoLcg rng((kPattern[layer] & st.bit) ? 5647381u : 7654321u);
            for (int k = 0; k < 2; ++k) {
                const int o = st.o0 + k * st.ostep, om = st.osum - o;
                for (int i = 4; i < W - 4; ++i) {
                    const int im = W - 1 - i;
                    uint8_t *a = n + 4 * (size_t) (st.outer_is_row ? i + o * W : o + i * W);
                    uint8_t *b = n + 4 * (size_t) (st.outer_is_row ? im + om * W : om + im * W);
                    for (int c = 0; c < 4; ++c) a[c] = b[c] = rng.byte();
                }
            }
        }
        for (int v = 4; v < W - 4; ++v)
            for (int h = 4; h < W - 4; ++h)
                for (int c = 0; c < 4; ++c) n[4 * (size_t) (h + v * W) + c] = interior.byte();
    }
}

extern "C" int pl_ortho_noise_host(int W, uint8_t *out)
{
    if (!out || W < 12 || W > 1024) return pl_set_error(PL_ERR_ARG, "bad argument");
    pl_host_ortho_noise(W, out);
    return PL_OK;
}

extern "C" int pl_ortho_noise_init(pl_ctx *ctx, int W, uint8_t *host_out)
{
    if (!ctx) return pl_set_error(PL_ERR_ARG, "ctx is NULL");
    if (W < 12 || W > 1024 || (W - 4) % 8 != 0)
        return pl_set_error(PL_ERR_ARG, "ortho tile_w - 4 must be a multiple of 8, got tile_w = %d", W);
    PL_CUDA(cudaSetDevice(ctx->device));
    std::vector<uint8_t> layers((size_t) 6 * W * W * 4);
    pl_host_ortho_noise(W, layers.data());
    if (host_out) memcpy(host_out, layers.data(), layers.size());
    /* rotation R reads texel (sel[R], sel[(R+1)%4]) with sel = (x, y, W-1-x, W-1-y)
     * (uvs[noiseUVLH.x], uvs[noiseUVLH.y] of the shader, NEAREST) */
    std::vector<uint32_t> rot((size_t) 24 * W * W);
    const uint32_t *src = (const uint32_t *) layers.data();
    for (int R = 0; R < 4; ++R)
        for (int L = 0; L < 6; ++L)
            for (int y = 0; y < W; ++y)
                for (int x = 0; x < W; ++x) {
                    const int sel[4] = { x, y, W - 1 - x, W - 1 - y };
                    rot[((size_t) (R * 6 + L) * W + y) * W + x] = src[((size_t) L * W + sel[(R + 1) % 4]) * W + sel[R]];
                }
    if (ctx->ortho_noise_rot) {
        PL_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->ortho_noise_rot);
        ctx->ortho_noise_rot = nullptr;
    }
    PL_CUDA(cudaMalloc(&ctx->ortho_noise_rot, rot.size() * 4));
    PL_CUDA(cudaMemcpyAsync(ctx->ortho_noise_rot, rot.data(), rot.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    PL_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->ortho_noise_w = W;
    return PL_OK;
}

/* --------------------------------------------------------------- host: requests */

extern "C" void pl_ortho_make_req(const pl_ortho_scene *sc, int level, int tx, int ty, int has_resid, pl_ortho_req *req)
{
    (void) has_resid;   /* the caller fills in the residual tile's slot */
    ortho_fill_req(pl_host_perlin(), sc, level, tx, ty, req);
}

extern "C" int pl_ortho_make_requests_range(const pl_ortho_scene *sc, int level, uint64_t morton0, int n, int out_slot0,
                                            int parent_slot0, uint64_t parent_morton0, pl_ortho_req *reqs, int nthreads)
{
    if (!sc || !reqs || n < 0 || level < 0 || level > 24) return pl_set_error(PL_ERR_ARG, "bad argument");
    const PerlinView T = pl_host_perlin();
    auto one = [&](int i) {
        const uint64_t m = morton0 + (uint64_t) i;
        int tx, ty;
        morton_decode(m, &tx, &ty);
        pl_ortho_req *q = reqs + i;
        ortho_fill_req(T, sc, level, tx, ty, q);
        q->out_slot = out_slot0 + i;
        q->parent_slot = level > 0 ? parent_slot0 + (int) ((m >> 2) - parent_morton0) : -1;
    };
    if (nthreads <= 0) nthreads = (int) std::thread::hardware_concurrency();
    if (nthreads < 1) nthreads = 1;
    if (n < 4096 || nthreads == 1) {
        for (int i = 0; i < n; ++i) one(i);
        return PL_OK;
    }
    std::vector<std::thread> pool;
    const int chunk = (n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; ++t) {
        const int lo = t * chunk, hi = lo + chunk < n ? lo + chunk : n;
        if (lo >= hi) break;
        pool.emplace_back([&one, lo, hi]() { for (int i = lo; i < hi; ++i) one(i); });
    }
    for (auto &th : pool) th.join();
    return PL_OK;
}

/* ----------------------------------------------------------------------- kernel */

namespace {

struct OrthoArgs {
    const pl_ortho_req *reqs;
    uint8_t *ortho;
    const uint8_t *resid;
    const uint32_t *noise_rot;
    long long slot_bytes, resid_slot_bytes;
    int W, PW;            /* tile width; parent window width W/2 + 2 */
    int channels;
    float scale;
    float root255[4];     /* rootNoiseColor * 255 */
};

constexpr int kOrthoThreads = 256;

__device__ __forceinline__ uint32_t ob_smem(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void ob_mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ob_smem(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void ob_mbar_expect(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ob_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ob_mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "OB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra OB_DONE;\n"
        "bra OB_WAIT;\n"
        "OB_DONE:\n"
        "}\n" ::"r"(ob_smem(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void ob_bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(ob_smem(dst)), "l"(src), "r"(bytes), "r"(ob_smem(bar)) : "memory");
}

/* RN(a / 255), RN(a / 6): q = a*y, r = a - b*q (exact), q' = q + r*y with y = RN(1/b) */
__device__ __forceinline__ float div255(float a) { return plfp::div_rn(a, 255.0f, 1.0f / 255.0f); }
__device__ __forceinline__ float div6(float a) { return plfp::div_rn(a, 6.0f, 1.0f / 6.0f); }
/* min(max(v, 0), 1) in one instruction (NaN -> 0 like fmaxf(NaN, 0)) */
__device__ __forceinline__ float clamp01(float v) { return __saturatef(v); }
__device__ __forceinline__ float fma_sat(float a, float b, float c)
{
    float d;
    asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

/* (float) byte K of w, minus BIAS, exactly and without the conversion unit: PRMT builds the float
 * 2^23 + b (0x4B0000bb), one FADD removes 2^23 + BIAS */
template <int K, int BIAS>
__device__ __forceinline__ float byte_f(uint32_t w)
{
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540u | K)) - (8388608.0f + (float) BIAS);
}

/* The RGBA8 colour buffer write of result / 255: clamp(RN(result / 255), 0, 1) * 255 rounded to nearest
 * even.  The rounding is the FADD of 1.5 * 2^23 (the sum's ulp is 1): the byte is the low byte of the
 * returned bits. */
__device__ __forceinline__ uint32_t to_unorm8_bits(float result)
{
    const float q = result * (1.0f / 255.0f);
    const float rem = fmaf(q, -255.0f, result);
    const float f = fma_sat(1.0f / 255.0f, rem, q);
    return __float_as_uint(f * 255.0f + 12582912.0f);
}
/* low bytes of four words -> one RGBA8 texel */
__device__ __forceinline__ uint32_t pack4(uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3)
{
    return __byte_perm(__byte_perm(b0, b1, 0x0040u), __byte_perm(b2, b3, 0x0040u), 0x5410u);
}

/* upsampleOrthoShader.glsl:141-151, the hsv branch: modulates r[0..2] in HSV space, r[3] directly.
 * Branch-free: neighbouring texels take different arms of rgb_to_hsv / hsv_to_rgb, so every arm is
 * evaluated and the result selected -- operation for operation what the taken arm of the oracle computes.
 *  - delta == 0: the divisions run on a substitute divisor 1 and H, S are selected to 0
 *  - S == 0 needs no arm of its own: V*(1 - 0), V*fma(-0, f, 1) are V exactly
 * MAYBE_NEG (residual variants): maxVal <= 0 with delta != 0 is possible (a residual below -parent) and
 * leaves the domain of the branch-free division: those texels take the IEEE operator. */
template <bool MAYBE_NEG>
__device__ __forceinline__ void hsv_noise(float r[4], const float nc[4], const float nm[4] /* noise - 128 */)
{
    const float R = div255(r[0]), G = div255(r[1]), B = div255(r[2]);
    const float minv = fminf(R, fminf(G, B)), maxv = fmaxf(R, fmaxf(G, B));
    const float delta = maxv - minv;
    const bool grey = delta == 0.0f;
    const float dsafe = grey ? 1.0f : delta;
    float msafe = maxv;
    if (MAYBE_NEG) msafe = maxv > 1e-30f ? maxv : 1.0f;
    else msafe = grey ? 1.0f : maxv;             /* without residuals rgb >= 0: delta != 0 implies maxv > 0 */
    float S = plfp::div_rn(delta, msafe, plfp::rcp_rn(msafe));
    if (MAYBE_NEG) {
        if (!(maxv > 1e-30f)) S = __fdiv_rn(delta, maxv);
    }
    const float rd = plfp::rcp_rn(dsafe);
    const float half = delta * 0.5f;                       /* delta / 2.0, exact */
    /* H = offset + del[a] - del[b] with (offset, a, b) = (0, B, G) when R is the maximum, (1/3, R, B) when G is,
     * (2/3, G, R) otherwise: only two of the shader's three del terms are ever used, so the two channels are
     * selected first and their terms evaluated once (0 + x is exact: the R branch's "del.z - del.y" is unchanged) */
    const bool rmax = R == maxv, gmax = G == maxv;
    const float ca = rmax ? B : (gmax ? R : G);
    const float cb = rmax ? G : (gmax ? B : R);
    const float hoff = rmax ? 0.0f : (gmax ? (float) (1.0 / 3.0) : (float) (2.0 / 3.0));
    const float da = plfp::div_rn(div6(maxv - ca) + half, dsafe, rd);
    const float db = plfp::div_rn(div6(maxv - cb) + half, dsafe, rd);
    float H = hoff + da - db;
    H = H < 0.0f ? H + 1.0f : H;
    H = H > 1.0f ? H - 1.0f : H;
    H = grey ? 0.0f : H;
    S = grey ? 0.0f : S;
    float V = maxv;
    constexpr float kEdge = 0.8f - 0.4f;
    const float e0 = V - 0.4f, tq = e0 * (1.0f / kEdge);
    const float t = fma_sat(1.0f / kEdge, fmaf(tq, -kEdge, e0), tq);     /* clamp(RN(e0 / kEdge), 0, 1) */
    const float k = 1.0f - t * t * fmaf(-2.0f, t, 3.0f);
    H *= 1.0f + div255(k * nc[0] * nm[0]);
    S *= 1.0f + div255(k * nc[1] * nm[1]);
    V *= 1.0f + div255(k * nc[2] * nm[2]);
    H = H - floorf(H);
    S = clamp01(S);
    V = clamp01(V);
    const float vh = H * 6.0f;
    const float vi = floorf(vh);
    const float f = vh - vi;
    const float v1 = V * (1.0f - S);
    const float v2 = V * fmaf(-S, f, 1.0f);
    const float v3 = V * fmaf(-S, 1.0f - f, 1.0f);
    /* sector:  0        1        2        3        4        other
     *   R      V        v2       v1       v1       v3       V
     *   G      v3       V        V        v2       v1       v1
     *   B      v1       v1       v3       V        V        v2     */
    const bool s0 = vi == 0.0f, s1 = vi == 1.0f, s2 = vi == 2.0f, s3 = vi == 3.0f, s4 = vi == 4.0f;
    const float oR = s1 ? v2 : ((s2 || s3) ? v1 : (s4 ? v3 : V));
    const float oG = s0 ? v3 : ((s1 || s2) ? V : (s3 ? v2 : v1));
    const float oB = (s0 || s1) ? v1 : (s2 ? v3 : ((s3 || s4) ? V : v2));
    r[0] = oR * 255.0f;
    r[1] = oG * 255.0f;
    r[2] = oB * 255.0f;
    r[3] = fmaf(nc[3], nm[3], r[3]);
}

/* all rows of one tile.  PARENT = false only for level-0 tiles */
template <bool HSV, bool RESID, bool PARENT, bool ALPHA>
__device__ __forceinline__ void ortho_rows(const OrthoArgs &a, const uint32_t *win, const uint32_t *noise, const uint8_t *res,
                                           uint8_t *out, const float nc[4], int tid)
{
    const int W = a.W, PW = a.PW;
    /* item i = y * groups + k: four texels from column 4k of row y.  i advances by the CTA size, so
     * (y, k) advance by its quotient and remainder with one conditional carry (no division in the loop) */
    const int groups = W >> 2;
    const int dy = kOrthoThreads / groups, dk = kOrthoThreads - dy * groups;
    int y = tid / groups, k = tid - y * groups;
    for (; y < W; y += dy, k += dk) {
        if (k >= groups) {
            k -= groups;
            if (++y >= W) break;
        }
        const size_t texel = (size_t) y * W + 4 * k;
        /* channels (0,2) in bytes 0 and 2 of cE, channels (1,3) in bytes 0 and 2 of cO (the other bytes are junk) */
        uint32_t cE[4] = { 0, 0, 0, 0 }, cO[4] = { 0, 0, 0, 0 };
        if (PARENT) {
            const uint32_t *r0 = win + ((y + 1) >> 1) * PW + 2 * k;
            const uint2 a01 = *(const uint2 *) r0, a23 = *(const uint2 *) (r0 + 2);
            const uint2 b01 = *(const uint2 *) (r0 + PW), b23 = *(const uint2 *) (r0 + PW + 2);
            const uint32_t wy0 = (y & 1) ? 3u : 1u, wy1 = 4u - wy0;
            const uint32_t ta[4] = { a01.x, a01.y, a23.x, a23.y }, tb[4] = { b01.x, b01.y, b23.x, b23.y };
            uint32_t vE[4], vO[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                /* two 16-bit lanes per register: bytes (0, 2) and (1, 3) of the texel */
                vE[j] = wy0 * __byte_perm(ta[j], 0u, 0x4240u) + wy1 * __byte_perm(tb[j], 0u, 0x4240u);
                vO[j] = wy0 * __byte_perm(ta[j], 0u, 0x4341u) + wy1 * __byte_perm(tb[j], 0u, 0x4341u);
            }
            /* texel 4k+p reads window columns 2k + (p+1)/2 and the next one; even x: weights (1,3), odd x: (3,1).
             * A lane's sum is < 4096, so after >> 4 its byte is clean and only bits 12..15 hold the neighbour's */
            cE[0] = (vE[0] + 3u * vE[1]) >> 4;  cO[0] = (vO[0] + 3u * vO[1]) >> 4;
            cE[1] = (3u * vE[1] + vE[2]) >> 4;  cO[1] = (3u * vO[1] + vO[2]) >> 4;
            cE[2] = (vE[1] + 3u * vE[2]) >> 4;  cO[2] = (vO[1] + 3u * vO[2]) >> 4;
            cE[3] = (3u * vE[2] + vE[3]) >> 4;  cO[3] = (3u * vO[2] + vO[3]) >> 4;
        }
        const uint4 nz4 = __ldg((const uint4 *) (noise + texel));
        const uint32_t nzw[4] = { nz4.x, nz4.y, nz4.z, nz4.w };
        uint32_t rw[4] = { 0, 0, 0, 0 };
        if (RESID && res) {
            if (a.channels == 4) {
                const uint4 r4 = __ldg((const uint4 *) (res + texel * 4));
                rw[0] = r4.x; rw[1] = r4.y; rw[2] = r4.z; rw[3] = r4.w;
            } else {
                const uint4 r4 = __ldg((const uint4 *) (res + texel * 4));
                /* a missing channel reads 0, a missing alpha 255 */
                const uint32_t keep = (1u << (8 * a.channels)) - 1u;
                rw[0] = (r4.x & keep) | 0xFF000000u; rw[1] = (r4.y & keep) | 0xFF000000u;
                rw[2] = (r4.z & keep) | 0xFF000000u; rw[3] = (r4.w & keep) | 0xFF000000u;
            }
        }
        uint32_t ow[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const float c[4] = { byte_f<0, 0>(cE[p]), byte_f<0, 0>(cO[p]), byte_f<2, 0>(cE[p]), byte_f<2, 0>(cO[p]) };
            /* noise - 128, residual - 128 */
            const float nm[4] = { byte_f<0, 128>(nzw[p]), byte_f<1, 128>(nzw[p]), byte_f<2, 128>(nzw[p]), byte_f<3, 128>(nzw[p]) };
            float r[4];
            if (RESID && res) {
                if (PARENT) {
                    r[0] = fmaf(byte_f<0, 128>(rw[p]), a.scale, c[0]);
                    r[1] = fmaf(byte_f<1, 128>(rw[p]), a.scale, c[1]);
                    r[2] = fmaf(byte_f<2, 128>(rw[p]), a.scale, c[2]);
                    r[3] = fmaf(byte_f<3, 128>(rw[p]), a.scale, c[3]);
                } else {
                    r[0] = byte_f<0, 0>(rw[p]); r[1] = byte_f<1, 0>(rw[p]); r[2] = byte_f<2, 0>(rw[p]); r[3] = byte_f<3, 0>(rw[p]);
                }
            } else {
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) r[ch] = PARENT ? c[ch] : a.root255[ch];
            }
            if (HSV) {
                hsv_noise<RESID>(r, nc, nm);
            } else {
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) r[ch] = fmaf(nc[ch], nm[ch], r[ch]);
            }
            if (ALPHA) {
                ow[p] = pack4(to_unorm8_bits(r[0]), to_unorm8_bits(r[1]), to_unorm8_bits(r[2]), to_unorm8_bits(r[3]));
            } else {   /* the storage keeps no alpha (RGB8, RG8, R8): its maths is dead code here, the byte is 0 */
                ow[p] = pack4(to_unorm8_bits(r[0]), to_unorm8_bits(r[1]), to_unorm8_bits(r[2]), 0u) & 0x00FFFFFFu;
            }
        }
        *(uint4 *) (out + texel * 4) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
}

template <bool HSV, bool RESID, bool ALPHA>
__global__ void __launch_bounds__(kOrthoThreads) ortho_kernel(const OrthoArgs a)
{
    extern __shared__ __align__(128) uint8_t ortho_smem[];
    __shared__ uint64_t bar;
    uint32_t *win = (uint32_t *) ortho_smem;

    const pl_ortho_req *qp = a.reqs + blockIdx.x;
    const int out_slot = qp->out_slot, parent_slot = qp->parent_slot;
    const int resid_slot = RESID ? qp->resid_slot : -1;
    const int W = a.W, PW = a.PW;
    const bool has_parent = parent_slot >= 0;
    const int tid = threadIdx.x;

    if (tid == 0) {
        ob_mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (has_parent && tid < 32) {
        if (tid == 0) ob_mbar_expect(&bar, (uint32_t) (PW * PW * 4));
        __syncwarp();
        const uint8_t *src = a.ortho + (long long) parent_slot * a.slot_bytes + ((size_t) qp->dy * W + qp->dx) * 4;
        for (int r = tid; r < PW; r += 32) ob_bulk_load(win + r * PW, src + (size_t) r * W * 4, (uint32_t) (PW * 4), &bar);
    }
    float nc[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) nc[c] = qp->noise_color[c];
    const uint32_t *noise = a.noise_rot + (size_t) (qp->noise_r * 6 + qp->noise_l) * W * W;
    uint8_t *out = a.ortho + (long long) out_slot * a.slot_bytes;
    const uint8_t *res = resid_slot >= 0 ? a.resid + (long long) resid_slot * a.resid_slot_bytes : nullptr;
    if (has_parent) ob_mbar_wait(&bar, 0);

    if (has_parent) ortho_rows<HSV, RESID, true, ALPHA>(a, win, noise, res, out, nc, tid);
    else ortho_rows<HSV, RESID, false, ALPHA>(a, win, noise, res, out, nc, tid);
}

}  // namespace

int pl_launch_ortho(pl_ctx *ctx, const pl_ortho_scene *sc, pl_pool *ortho, pl_pool *resid, int n,
                    const pl_ortho_req *dev_reqs)
{
    PL_CUDA(cudaSetDevice(ctx->device));
    OrthoArgs a;
    a.reqs = dev_reqs;
    a.ortho = ortho->base;
    a.resid = resid ? resid->base : nullptr;
    a.noise_rot = ctx->ortho_noise_rot;
    a.slot_bytes = (long long) ortho->slot_bytes;
    a.resid_slot_bytes = resid ? (long long) resid->slot_bytes : 0;
    a.W = sc->tile_w;
    a.PW = sc->tile_w / 2 + 2;
    a.channels = sc->channels;
    a.scale = sc->scale;
    for (int c = 0; c < 4; ++c) a.root255[c] = sc->root_noise_color[c] * 255.0f;
    const size_t smem = (size_t) a.PW * a.PW * 4;
    if (smem > 227 * 1024) return pl_set_error(PL_ERR_ARG, "ortho tile_w %d needs %zu bytes of shared memory", a.W, smem);
    /* out_channels 1..3: the storage has no alpha channel (RGB8 in terrain3/helloworld.xml:43): it is not computed */
    const bool alpha = !(sc->out_channels >= 1 && sc->out_channels <= 3);
    void (*kern)(const OrthoArgs) =
        alpha ? (sc->hsv ? (resid ? ortho_kernel<true, true, true> : ortho_kernel<true, false, true>)
                         : (resid ? ortho_kernel<false, true, true> : ortho_kernel<false, false, true>))
              : (sc->hsv ? (resid ? ortho_kernel<true, true, false> : ortho_kernel<true, false, false>)
                         : (resid ? ortho_kernel<false, true, false> : ortho_kernel<false, false, false>));
    if (smem > 40 * 1024) PL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    pl_timing_begin(ctx, PL_K_ORTHO, n);
    kern<<<n, kOrthoThreads, smem, ctx->stream>>>(a);
    pl_timing_end(ctx);
    PL_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return PL_OK;
}

/* ------------------------------------------------------------------ entry points */

static int check_ortho_args(pl_ctx *ctx, const pl_ortho_scene *sc, const pl_pool *ortho, const pl_pool *resid, int n)
{
    if (!ctx || !sc || !ortho) return pl_set_error(PL_ERR_ARG, "NULL argument");
    if (n < 0) return pl_set_error(PL_ERR_ARG, "n = %d", n);
    /* an RGBA8 normal pool has the same layout (dense 4-byte texels): a host layer that creates its
     * RGBA8 storages before knowing their producer passes one */
    auto rgba8 = [](const pl_pool *p) {
        return (p->kind == PL_POOL_ORTHO_UN8x4 || p->kind == PL_POOL_NORM_UN8x4) && (p->tile_w - 4) % 8 == 0;
    };
    if (!rgba8(ortho) || ortho->ctx != ctx || ortho->tile_w != sc->tile_w)
        return pl_set_error(PL_ERR_ARG, "ortho pool: wrong kind, context or tile_w (scene %d)", sc->tile_w);
    if (resid && (!rgba8(resid) || resid->ctx != ctx || resid->tile_w != sc->tile_w))
        return pl_set_error(PL_ERR_ARG, "ortho residual pool: wrong kind, context or tile_w");
    if (sc->channels < 1 || sc->channels > 4) return pl_set_error(PL_ERR_ARG, "channels = %d", sc->channels);
    if (!ctx->ortho_noise_rot || ctx->ortho_noise_w != sc->tile_w)
        return pl_set_error(PL_ERR_ARG, "pl_ortho_noise_init(%d) has not been called", sc->tile_w);
    return PL_OK;
}

extern "C" int pl_ortho_batch_dev(pl_ctx *ctx, const pl_ortho_scene *sc, pl_pool *ortho, pl_pool *resid, int n,
                                  const pl_ortho_req *dev_reqs)
{
    int rc = check_ortho_args(ctx, sc, ortho, resid, n);
    if (rc) return rc;
    if (n == 0) return PL_OK;
    if (!dev_reqs) return pl_set_error(PL_ERR_ARG, "reqs is NULL");
    return pl_launch_ortho(ctx, sc, ortho, resid, n, dev_reqs);
}

extern "C" int pl_ortho_batch(pl_ctx *ctx, const pl_ortho_scene *sc, pl_pool *ortho, pl_pool *resid, int n,
                              const pl_ortho_req *reqs)
{
    int rc = check_ortho_args(ctx, sc, ortho, resid, n);
    if (rc) return rc;
    if (n == 0) return PL_OK;
    if (!reqs) return pl_set_error(PL_ERR_ARG, "reqs is NULL");
    const int half = (sc->tile_w - 4) / 2;
    for (int i = 0; i < n; ++i) {
        const pl_ortho_req &q = reqs[i];
        if (q.out_slot < 0 || q.out_slot >= ortho->capacity || q.parent_slot >= ortho->capacity ||
            (q.parent_slot >= 0 && q.parent_slot == q.out_slot))
            return pl_set_error(PL_ERR_ARG, "request %d: slot out of range", i);
        if (q.resid_slot >= 0 && (!resid || q.resid_slot >= resid->capacity))
            return pl_set_error(PL_ERR_ARG, "request %d: residual slot %d without a pool / out of range", i, q.resid_slot);
        if (q.parent_slot >= 0 && ((q.dx != 0 && q.dx != half) || (q.dy != 0 && q.dy != half)))
            return pl_set_error(PL_ERR_ARG, "request %d: dx, dy must be 0 or %d", i, half);
        if ((unsigned) q.noise_r > 3u || (unsigned) q.noise_l > 5u)
            return pl_set_error(PL_ERR_ARG, "request %d: noise rotation / layer out of range", i);
    }
    void *dev = nullptr;
    rc = pl_stage_requests(ctx, reqs, sizeof(pl_ortho_req) * (size_t) n, &dev);
    if (rc) return rc;
    return pl_launch_ortho(ctx, sc, ortho, resid, n, (const pl_ortho_req *) dev);
}
