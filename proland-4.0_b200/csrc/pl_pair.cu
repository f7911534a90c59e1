/*
 * pl_pair.cu -- the fused elevation + normal pass for sm_100a: one CTA produces the elevation tile
 * (zf, zc, zm planes) AND the RG8 normal tile derived from it.
 *
 * The reference runs two GL passes per tile pair: ElevationProducer::doCreateTile draws upsampleShader
 * (ElevationProducer.cpp:280-405, upsampleShader.glsl:140-203), then NormalProducer::doCreateTile
 * draws normalShader on the finished elevation tile (NormalProducer.cpp:164-289,
 * normalShader.glsl:60-125).  The two separate kernels (pl_elevation.cu, pl_normal.cu) mirror that; on
 * B200 they leave each other's resource idle: the elevation pass is bound by the HBM write of the three
 * planes and keeps the fp32 pipe about 60 % busy, the normal pass is bound by the fp32 pipe (the
 * spherical world position of every grid point) and uses about 15 % of the HBM bandwidth.  Fused,
 *   - the zm plane the normal pass needs never comes back from HBM: the elevation phase leaves a copy
 *     in shared memory (39 KB of the 58 KB the normal pass would read and write)
 *   - CTAs resident on one SM are in different phases, so the fp32 pipe works on normals while other
 *     CTAs wait for their elevation stores, and one launch replaces two
 * The per-tile device code is the same as the separate kernels' (pl_elevation_tile.cuh,
 * pl_normal_tile.cuh): results are bit-identical.
 *
 * Shared memory of a CTA: [guard 96 B | zm plane 42 016 B | elevation windows + tables, reused as the
 * position planes of the normal phase 27 456 B | uv tables 816 B | row table (PL_ARITH_FAST) 1 408 B] = 71.7 KB
 * -> 3 CTAs per SM.
 */
#include "pl_elevation_tile.cuh"
#include "pl_normal_tile.cuh"

int pl_elev_fill_args(pl_ctx *ctx, const pl_elev_scene *sc, pl_pool *elev, pl_pool *resid, const pl_elev_req *dev_reqs,
                      plelev::ElevArgs &a);
int pl_norm_fill_args(pl_ctx *ctx, const pl_norm_scene *sc, pl_pool *norm, pl_pool *elev, const pl_norm_req *dev_reqs,
                      plnorm::NormArgs &a);

namespace {


/* SLIM: the layout of launches in which EVERY tile takes the register form of the normal pass (PL_ARITH_FAST, the host
 * checks it per launch): no position planes, the elevation phase without its second window copy, 28 floats of corner
 * values instead of the banded form's row table -- 55.6 KB, FOUR CTAs per SM instead of three (71.7 KB). */
template <int TW, int TG, bool SLIM = false>
struct PairSmem {
    using EG = plelev::Geo<TW, TG>;
    using NG = plnorm::NGeo<TW - 4>;
    static constexpr int GUARD = 24;                                           /* floats in front of the zm plane (>= 4; makes ZM a multiple of 128 bytes) */
    static constexpr int ZM = GUARD + EG::PLANE;                               /* floats: guard + zm plane */
    static constexpr int ELEV = plelev::ElevSmem<TW, TG, !SLIM>::FLOATS;
    static constexpr int POS = SLIM ? 0 : 3 * NG::POS_PLANE;
    static constexpr int WORK = ((ELEV > POS ? ELEV : POS) + 3) & ~3;           /* elevation scratch, then positions */
    static constexpr int QTAB = SLIM ? 32 : NG::ROWTAB;
    static constexpr size_t BYTES = (size_t) (ZM + WORK + 2 * NG::ULUT + QTAB) * 4;
    static_assert(EG::PITCH == NG::EPITCH && EG::PLANE == NG::EPLANE, "both passes agree on the plane layout");
    static_assert((ZM * 4) % 128 == 0, "the TMA destination (first window) stays 128-byte aligned");
    static_assert(ELEV - 64 >= plnorm::RGeo<TW - 4, 224>::TAB && ELEV - 64 >= plnorm::RGeo<TW - 4, 192>::TAB && QTAB >= 28,
                  "row tables of the register form fit in front of the statistics");
    static_assert(!SLIM || BYTES + 1280 <= 233472 / 4, "four CTAs per SM");
};

/* threads of a CTA: 256 under PL_ARITH_EXACT; 224 under PL_ARITH_FAST -- 7 warps x 3 CTAs leave 96 registers per
 * thread, which the register form of the normal pass needs (three grid rows of positions live in registers), and the
 * 663 quad pairs of an elevation tile are 3 passes of 224 threads (97 % of the lanes busy) as they are 3 passes of 256 */
template <bool FAST> struct PairThreads { static constexpr int N = FAST ? 224 : 256; };

template <int TW, int TG, int RESID, bool SPHERE, bool LINEAR, bool PUSH = false, bool FAST = false, bool SLIM = false, bool C4 = false,
          int NTHREADS = PairThreads<FAST>::N>
__global__ void __launch_bounds__(NTHREADS, SLIM ? 4 : 3)
tile_pair_kernel(const __grid_constant__ CUtensorMap tm, const plelev::ElevArgs ea, const plnorm::NormArgs na)
{
    static_assert(!SLIM || (FAST && !PUSH), "the slim layout serves the register form only");
    static_assert(!C4 || (!FAST && !PUSH && !SLIM), "RGBA8 normals (fine + parent coarse normal): exact arithmetic, no push");
    using SM = PairSmem<TW, TG, SLIM>;
    constexpr int kPairThreads = NTHREADS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *zs = reinterpret_cast<float *>(smem_raw) + SM::GUARD;   /* zm plane behind the guard floats */
    float *work = zs + plelev::Geo<TW, TG>::PLANE;
    float *ulut = work + SM::WORK;
    float *rowtab = ulut + 2 * SM::NG::ULUT;   /* PL_ARITH_FAST: the row table of a band of the normal phase */
    __shared__ uint64_t bar;
    __shared__ pl_norm_req nrq;

    const int tid = threadIdx.x;
    const int tile = blockIdx.x;
    const pl_elev_req erq = ea.reqs[tile];
    {
        const int *src = reinterpret_cast<const int *>(na.reqs + tile);
        int *dst = reinterpret_cast<int *>(&nrq);
        if (tid < (int) (sizeof(pl_norm_req) / 4)) dst[tid] = __ldg(src + tid);
    }
    if (tid == 0) {
        plelev::mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    plnorm::normal_uv_tables<TW - 4, kPairThreads>(ulut, tid);
    __syncthreads();
    if (FAST && SPHERE) plnorm::normal_reg_qtab(rowtab, nrq, tid);   /* visible after the elevation phase's barriers */
    plelev::levels_wait_parent(ea, erq, tid);                        /* pl_produce_levels: the parent is a CTA of this launch */

    /* elevation: planes to HBM, zm also to shared memory */
    unsigned short *out = reinterpret_cast<unsigned short *>(na.norm + (size_t) nrq.out_slot * na.norm_slot_bytes);
    const bool reg_form = SLIM || (FAST && !PUSH && plnorm::normal_reg_ok(nrq, SPHERE));
    plelev::elevation_tile<TW, TG, RESID, kPairThreads, true, !SLIM>(&tm, ea, erq, work, &bar, 0, zs, tid, reg_form);
    plelev::levels_publish(ea, erq, tid);   /* children may start while this CTA computes its normals */
    __syncthreads();   /* zm plane complete; the elevation scratch is free */

    /* normals from the shared zm plane */
    if (reg_form) {
        /* the register form: no position planes, no barriers; its row tables live in the elevation scratch (the
         * first RGeo::TAB floats of it: the per-warp statistics at its end stay intact and are finished here,
         * behind the barrier above instead of one of their own) */
        if (tid == kPairThreads - 32 && ea.want_stats) plelev::elevation_stats_finish<TW, TG, kPairThreads, !SLIM>(ea, erq, work);
        plnorm::normal_tile_reg<TW - 4, SPHERE, LINEAR, kPairThreads>(zs, work, rowtab, ulut, nrq, out, tid);
        return;
    }
    if (!SLIM) plnorm::normal_tile<TW - 4, SPHERE, LINEAR, kPairThreads, PUSH, FAST, C4>(zs, work, ulut, nrq, out, tid, &na, rowtab);
}

template <int RESID>
int launch_pair(pl_ctx *ctx, pl_pool *elev, const plelev::ElevArgs &ea, const plnorm::NormArgs &na, int n, bool all_reg = false)
{
    using SM = PairSmem<101, 4>;
    if (all_reg && na.fast && na.npeers == 0 && na.channels != 4 && !ctx->no_slim && (!na.sphere || na.linear || ctx->slim_sphere)) {
        /* every tile of the launch qualifies for the register form (the caller checked): the slim layout, 4 CTAs per SM.
         * Flat scenes: 224 threads, 70 registers, + 6.6 % (0.583 -> 0.547 ms per 16 384 pairs).  Spheres: the register form
         * needs 80 registers, which 4 CTAs of 224 threads do not have (72: spills, no gain) -- 192 threads do (6 warps x 4 =
         * 24 warps per SM against 21): + 2.1 % (0.757 -> 0.742 ms) with LINEAR elevation storages; with NEAREST ones the
         * normal phase is lighter, the extra loads of the window-less elevation phase weigh more and the slim layout LOSES
         * 1.5 % (0.717 -> 0.728 ms): those keep the regular layout (tools/microbench/slim_compare.py; pl_debug_no_slim(ctx, -1)
         * forces the slim one) */
        using SL = PairSmem<101, 4, true>;
        /* spheres: 192 threads (6 warps x 4 CTAs leave the 80 registers the register form needs: no spills) */
        void (*slim)(const CUtensorMap, const plelev::ElevArgs, const plnorm::NormArgs) =
            na.sphere ? (na.linear ? tile_pair_kernel<101, 4, RESID, true, true, false, true, true, false, 192> : tile_pair_kernel<101, 4, RESID, true, false, false, true, true, false, 192>)
                      : (na.linear ? tile_pair_kernel<101, 4, RESID, false, true, false, true, true> : tile_pair_kernel<101, 4, RESID, false, false, false, true, true>);
        PL_CUDA(cudaFuncSetAttribute(slim, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) SL::BYTES));
        pl_timing_begin(ctx, PL_K_PAIR, n);
        slim<<<n, na.sphere ? 192 : PairThreads<true>::N, SL::BYTES, ctx->stream>>>(elev->tm_parent, ea, na);
        pl_timing_end(ctx);
        PL_CUDA(cudaGetLastError());
        ctx->launches += 1;
        return PL_OK;
    }
    void (*kern)(const CUtensorMap, const plelev::ElevArgs, const plnorm::NormArgs) =
        na.sphere ? (na.linear ? tile_pair_kernel<101, 4, RESID, true, true> : tile_pair_kernel<101, 4, RESID, true, false>)
                  : (na.linear ? tile_pair_kernel<101, 4, RESID, false, true> : tile_pair_kernel<101, 4, RESID, false, false>);
    if (na.fast && na.npeers == 0 && na.channels != 4)
        kern = na.sphere ? (na.linear ? tile_pair_kernel<101, 4, RESID, true, true, false, true> : tile_pair_kernel<101, 4, RESID, true, false, false, true>)
                         : (na.linear ? tile_pair_kernel<101, 4, RESID, false, true, false, true> : tile_pair_kernel<101, 4, RESID, false, false, false, true>);
    if (na.channels == 4) {
        /* RGBA8 normal storage: fine normal + the parent's coarse normal (normalShader.glsl:100-114); exact arithmetic */
        if (na.npeers > 0) return pl_set_error(PL_ERR_ARG, "pushing tiles to peers is built for RG8 normal pools");
        kern = na.sphere ? (na.linear ? tile_pair_kernel<101, 4, RESID, true, true, false, false, false, true> : tile_pair_kernel<101, 4, RESID, true, false, false, false, false, true>)
                         : (na.linear ? tile_pair_kernel<101, 4, RESID, false, true, false, false, false, true> : tile_pair_kernel<101, 4, RESID, false, false, false, false, false, true>);
    }
    if (na.npeers > 0) {
        /* finished normal tiles also go to the peer GPUs (fractal scenes: the variants without residuals) */
        if (RESID != 0) return pl_set_error(PL_ERR_ARG, "pushing tiles to peers is built for scenes without residuals");
        kern = na.sphere ? (na.linear ? tile_pair_kernel<101, 4, 0, true, true, true> : tile_pair_kernel<101, 4, 0, true, false, true>)
                         : (na.linear ? tile_pair_kernel<101, 4, 0, false, true, true> : tile_pair_kernel<101, 4, 0, false, false, true>);
    }
    PL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) SM::BYTES));
    pl_timing_begin(ctx, PL_K_PAIR, n);
    const int threads = na.fast && na.npeers == 0 && na.channels != 4 ? PairThreads<true>::N : PairThreads<false>::N;
    kern<<<n, threads, SM::BYTES, ctx->stream>>>(elev->tm_parent, ea, na);
    pl_timing_end(ctx);
    PL_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return PL_OK;
}

}  // namespace

/* true when the fused kernel serves this geometry (the one of every shipped archive) */
bool pl_pair_supported(const pl_ctx *ctx, const pl_elev_scene *esc, const pl_norm_scene *nsc, const pl_pool *elev,
                       const pl_pool *norm)
{
    return !ctx->force_generic && !ctx->no_fuse && elev->tile_w == 101 && esc->grid == 4 && norm->tile_w == 97 &&
           nsc->elev_border == 2 && (norm->kind == PL_POOL_NORM_UN8x2 || norm->kind == PL_POOL_NORM_UN8x4);
}

/* n tile pairs: normal request i must describe the normal tile of elevation request i
 * (nreq[i].elev_slot == ereq[i].out_slot) */
static int launch_pair_any(pl_ctx *ctx, const pl_elev_scene *esc, const pl_norm_scene *nsc, pl_pool *elev, pl_pool *norm,
                           pl_pool *resid, int n, const pl_elev_req *dev_ereqs, const pl_norm_req *dev_nreqs, int epoch, bool all_reg);

int pl_launch_pair(pl_ctx *ctx, const pl_elev_scene *esc, const pl_norm_scene *nsc, pl_pool *elev, pl_pool *norm,
                   pl_pool *resid, int n, const pl_elev_req *dev_ereqs, const pl_norm_req *dev_nreqs, bool all_reg)
{
    return launch_pair_any(ctx, esc, nsc, elev, norm, resid, n, dev_ereqs, dev_nreqs, 0, all_reg);
}

/* tiles of several levels in one launch, ordered by level: requests with pad_[0] != 0 wait for their parent's ready flag */
int pl_launch_pair_levels(pl_ctx *ctx, const pl_elev_scene *esc, const pl_norm_scene *nsc, pl_pool *elev, pl_pool *norm, int n,
                          const pl_elev_req *dev_ereqs, const pl_norm_req *dev_nreqs, int epoch, bool all_reg)
{
    return launch_pair_any(ctx, esc, nsc, elev, norm, nullptr, n, dev_ereqs, dev_nreqs, epoch, all_reg);
}

static int launch_pair_any(pl_ctx *ctx, const pl_elev_scene *esc, const pl_norm_scene *nsc, pl_pool *elev, pl_pool *norm,
                           pl_pool *resid, int n, const pl_elev_req *dev_ereqs, const pl_norm_req *dev_nreqs, int epoch, bool all_reg)
{
    PL_CUDA(cudaSetDevice(ctx->device));
    plelev::ElevArgs ea;
    plnorm::NormArgs na;
    int rc = pl_elev_fill_args(ctx, esc, elev, resid, dev_ereqs, ea);
    if (rc) return rc;
    if (epoch != 0) {
        ea.ready = elev->ready;
        ea.epoch = epoch;
    }
    if ((rc = pl_norm_fill_args(ctx, nsc, norm, elev, dev_nreqs, na)) != PL_OK) return rc;
    const int rk = !resid ? 0 : (resid->kind == PL_POOL_RESID_F32 ? 1 : 2);
    return rk == 0 ? launch_pair<0>(ctx, elev, ea, na, n, all_reg) : (rk == 1 ? launch_pair<1>(ctx, elev, ea, na, n, all_reg) : launch_pair<2>(ctx, elev, ea, na, n, all_reg));
}
