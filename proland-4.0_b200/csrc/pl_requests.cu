/*
 * pl_requests.cu -- the device-side batch driver: requests for a whole Morton
 * range of tiles are generated ON THE GPU (noise layer selection through cnoise,
 * fp64 patch geometry), then the elevation and normal kernels run on them.
 *
 * This is what replaces the reference's per-tile host loop
 * (TileProducer::createTile -> CreateTile::run -> doCreateTile, one GL draw per
 * tile: producer/TileProducer.cpp:199-217, ElevationProducer.cpp:280-405): the
 * host only says "level L, Morton range [m0, m0+n), slots from s0", 8 bytes of
 * arguments per batch instead of ~300 bytes of uniforms per tile.
 *
 * The maths is pl_reqmath.cuh, the same source the host entry points compile,
 * and is bit-identical to it (tests/test_gpu_parity.py::test_device_requests).
 */
#include <cstring>
#include <thread>
#include <vector>

#include "pl_reqmath.cuh"

PerlinView pl_host_perlin();

namespace {

struct GenArgs {
    PerlinView perlin;
    pl_elev_req *ereq;
    pl_norm_req *nreq;      /* may be NULL */
    int tile_w, face, n_amp, sphere;
    float root_quad_size;
    float noise_amp[32];
    int level, n;
    unsigned long long morton0, parent_morton0;
    int out_slot0, parent_slot0;
    int norm4;              /* RGBA8 normal pool: the normal request names the parent's normal tile (same slot numbering) */
};

__host__ __device__ __forceinline__ void gen_one(const GenArgs &g, int i)
{
    const unsigned long long m = g.morton0 + (unsigned long long) i;
    int tx, ty;
    morton_decode(m, &tx, &ty);
    pl_elev_req e;
    elev_fill_req(g.perlin, g.tile_w, g.root_quad_size, g.noise_amp, g.n_amp, g.face, g.level, tx, ty, 0, 0, &e);
    e.out_slot = g.out_slot0 + i;
    e.parent_slot = g.level > 0 ? g.parent_slot0 + (int) ((m >> 2) - g.parent_morton0) : -1;
    g.ereq[i] = e;
    if (g.nreq) {
        pl_norm_req q;
        norm_fill_req(g.sphere, (double) g.root_quad_size, g.level, tx, ty, &q);
        q.out_slot = g.out_slot0 + i;
        q.elev_slot = g.out_slot0 + i;
        if (g.norm4) q.parent_slot = e.parent_slot;      /* normalOSL: the parent's normal tile sits in the parent's slot */
        g.nreq[i] = q;
    }
}

__global__ void __launch_bounds__(128) gen_requests_kernel(const GenArgs g)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < g.n) gen_one(g, i);
}

int ensure_perlin(pl_ctx *ctx)
{
    if (ctx->perlin_perm) return PL_OK;
    const PerlinView h = pl_host_perlin();
    PL_CUDA(cudaMalloc(&ctx->perlin_perm, sizeof(int) * kPerlinN));
    PL_CUDA(cudaMalloc(&ctx->perlin_g2, sizeof(float) * 2 * kPerlinN));
    PL_CUDA(cudaMemcpyAsync(ctx->perlin_perm, h.perm, sizeof(int) * kPerlinN, cudaMemcpyHostToDevice, ctx->stream));
    PL_CUDA(cudaMemcpyAsync(ctx->perlin_g2, h.g2, sizeof(float) * 2 * kPerlinN, cudaMemcpyHostToDevice, ctx->stream));
    return PL_OK;
}

int ensure_gen_buffers(pl_ctx *ctx, int n)
{
    if (n <= ctx->gen_cap) return PL_OK;
    PL_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->gen_ereq) cudaFree(ctx->gen_ereq);
    if (ctx->gen_nreq) cudaFree(ctx->gen_nreq);
    ctx->gen_ereq = nullptr;
    ctx->gen_nreq = nullptr;
    ctx->gen_cap = 0;
    const int cap = n + n / 4 + 256;
    PL_CUDA(cudaMalloc(&ctx->gen_ereq, sizeof(pl_elev_req) * (size_t) cap));
    PL_CUDA(cudaMalloc(&ctx->gen_nreq, sizeof(pl_norm_req) * (size_t) cap));
    ctx->gen_cap = cap;
    return PL_OK;
}

int check_range(const pl_sweep_scene *sc, pl_pool *elev, pl_pool *norm, int level, uint64_t morton0, int n,
                int out_slot0, int parent_slot0, uint64_t parent_morton0)
{
    if (!sc || !elev) return pl_set_error(PL_ERR_ARG, "NULL argument");
    if (level < 0 || level > 24) return pl_set_error(PL_ERR_ARG, "level %d out of range", level);
    const uint64_t count = 1ull << (2 * level);
    if (n < 0 || morton0 + (uint64_t) n > count) return pl_set_error(PL_ERR_ARG, "Morton range exceeds level %d", level);
    if (out_slot0 < 0 || out_slot0 + n > elev->capacity || (norm && out_slot0 + n > norm->capacity))
        return pl_set_error(PL_ERR_POOL_FULL, "slots [%d,%d) exceed the pool capacity", out_slot0, out_slot0 + n);
    if (level > 0 && n > 0) {
        const uint64_t p_first = morton0 >> 2, p_last = (morton0 + n - 1) >> 2;
        if (p_first < parent_morton0) return pl_set_error(PL_ERR_ARG, "parent range starts after the first parent");
        const int64_t s_first = (int64_t) parent_slot0 + (int64_t) (p_first - parent_morton0);
        const int64_t s_last = (int64_t) parent_slot0 + (int64_t) (p_last - parent_morton0);
        if (s_first < 0 || s_last >= elev->capacity) return pl_set_error(PL_ERR_ARG, "parent slots out of range");
        if (s_first < (int64_t) out_slot0 + n && s_last >= out_slot0)
            return pl_set_error(PL_ERR_ARG, "parent slots overlap the output slots");
    }
    if (sc->n_amp < 0 || sc->n_amp > 32) return pl_set_error(PL_ERR_ARG, "n_amp out of range");
    return PL_OK;
}

void fill_gen_args(GenArgs &g, const pl_sweep_scene *sc, int level, uint64_t morton0, int n, int out_slot0,
                   int parent_slot0, uint64_t parent_morton0)
{
    g.tile_w = sc->elev.tile_w;
    g.face = sc->face;
    g.n_amp = sc->n_amp;
    g.sphere = sc->norm.sphere;
    g.root_quad_size = sc->root_quad_size;
    memcpy(g.noise_amp, sc->noise_amp, sizeof(g.noise_amp));
    g.level = level;
    g.n = n;
    g.morton0 = morton0;
    g.parent_morton0 = parent_morton0;
    g.out_slot0 = out_slot0;
    g.parent_slot0 = parent_slot0;
    g.norm4 = 0;
}

}  // namespace

extern "C" int pl_produce_range(pl_ctx *ctx, const pl_sweep_scene *sc, pl_pool *elev, pl_pool *norm, int level,
                                uint64_t morton0, int n, int out_slot0, int parent_slot0, uint64_t parent_morton0)
{
    if (!ctx) return pl_set_error(PL_ERR_ARG, "ctx is NULL");
    int rc = check_range(sc, elev, norm, level, morton0, n, out_slot0, parent_slot0, parent_morton0);
    if (rc) return rc;
    if (n == 0) return PL_OK;
    PL_CUDA(cudaSetDevice(ctx->device));
    if ((rc = ensure_perlin(ctx)) != PL_OK) return rc;
    if ((rc = ensure_gen_buffers(ctx, n)) != PL_OK) return rc;
    GenArgs g;
    g.perlin.perm = ctx->perlin_perm;
    g.perlin.g2 = ctx->perlin_g2;
    g.ereq = ctx->gen_ereq;
    g.nreq = norm ? ctx->gen_nreq : nullptr;
    fill_gen_args(g, sc, level, morton0, n, out_slot0, parent_slot0, parent_morton0);
    g.norm4 = norm && norm->kind == PL_POOL_NORM_UN8x4;
    pl_timing_begin(ctx, PL_K_GENREQ, n);
    gen_requests_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(g);
    pl_timing_end(ctx);
    PL_CUDA(cudaGetLastError());
    ctx->launches += 1;
    /* request i of both arrays describes tile i: the fused kernel applies */
    if (norm && pl_pair_supported(ctx, &sc->elev, &sc->norm, elev, norm))
        return pl_launch_pair(ctx, &sc->elev, &sc->norm, elev, norm, nullptr, n, ctx->gen_ereq, ctx->gen_nreq,
                              norm_level_all_reg(sc->norm.sphere, sc->root_quad_size, level));
    rc = pl_elevation_batch_dev(ctx, &sc->elev, elev, nullptr, n, ctx->gen_ereq);
    if (rc) return rc;
    if (norm) rc = pl_normal_batch_dev(ctx, &sc->norm, norm, elev, n, ctx->gen_nreq);
    return rc;
}

namespace {

struct GenIdArgs {
    PerlinView perlin;
    pl_elev_req *ereq;
    pl_norm_req *nreq;
    const pl_tile_id *ids;
    int tile_w, face, n_amp, sphere, resid_tile_w, n;
    float root_quad_size;
    float noise_amp[32];
};

__global__ void __launch_bounds__(128) gen_requests_ids_kernel(const GenIdArgs g)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    const pl_tile_id t = g.ids[i];
    pl_elev_req e;
    elev_fill_req(g.perlin, g.tile_w, g.root_quad_size, g.noise_amp, g.n_amp, g.face, t.level, t.tx, t.ty, g.resid_tile_w,
                  t.resid_slot >= 0, &e);
    e.out_slot = t.elev_slot;
    e.parent_slot = t.level > 0 ? t.parent_slot : -1;
    e.resid_slot = t.resid_slot >= 0 ? t.resid_slot : -1;
    g.ereq[i] = e;
    pl_norm_req q;
    norm_fill_req(g.sphere, (double) g.root_quad_size, t.level, t.tx, t.ty, &q);
    q.out_slot = t.norm_slot;
    q.elev_slot = t.elev_slot;
    g.nreq[i] = q;
}

}  // namespace

extern "C" int pl_pair_batch_ids(pl_ctx *ctx, const pl_sweep_scene *sc, pl_pool *elev, pl_pool *norm, pl_pool *resid, int n,
                                 const pl_tile_id *ids)
{
    if (!ctx || !sc || !elev || !norm) return pl_set_error(PL_ERR_ARG, "NULL argument");
    if (n < 0) return pl_set_error(PL_ERR_ARG, "n < 0");
    if (n == 0) return PL_OK;
    if (!ids) return pl_set_error(PL_ERR_ARG, "ids is NULL");
    if (sc->n_amp < 0 || sc->n_amp > 32) return pl_set_error(PL_ERR_ARG, "n_amp out of range");
    if (elev->kind != PL_POOL_ELEV_F32x3 || (norm->kind != PL_POOL_NORM_UN8x2 && norm->kind != PL_POOL_NORM_UN8x4))
        return pl_set_error(PL_ERR_ARG, "pl_pair_batch_ids needs an elevation and a normal pool");
    if (resid && resid->kind != PL_POOL_RESID_F32 && resid->kind != PL_POOL_RESID_I16)
        return pl_set_error(PL_ERR_ARG, "resid is not a residual pool");
    const int rcap = resid ? resid->capacity : 0;
    int min_level = 99;
    for (int i = 0; i < n; ++i) {
        const pl_tile_id &t = ids[i];
        min_level = t.level < min_level ? t.level : min_level;
        if (t.level < 0 || t.level > 24 || t.tx < 0 || t.ty < 0 || t.tx >= (1 << t.level) || t.ty >= (1 << t.level))
            return pl_set_error(PL_ERR_ARG, "tile %d: (%d, %d, %d) is not a quadtree tile", i, t.level, t.tx, t.ty);
        if (t.elev_slot < 0 || t.elev_slot >= elev->capacity || t.norm_slot < 0 || t.norm_slot >= norm->capacity ||
            t.resid_slot >= rcap || (t.level > 0 && (t.parent_slot < 0 || t.parent_slot >= elev->capacity || t.parent_slot == t.elev_slot)))
            return pl_set_error(PL_ERR_ARG, "tile %d: slot out of range", i);
    }
    PL_CUDA(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ensure_perlin(ctx)) != PL_OK) return rc;
    if ((rc = ensure_gen_buffers(ctx, n)) != PL_OK) return rc;
    void *dev = nullptr;
    if ((rc = pl_stage_requests(ctx, ids, sizeof(pl_tile_id) * (size_t) n, &dev)) != PL_OK) return rc;
    GenIdArgs g;
    g.perlin.perm = ctx->perlin_perm;
    g.perlin.g2 = ctx->perlin_g2;
    g.ereq = ctx->gen_ereq;
    g.nreq = ctx->gen_nreq;
    g.ids = static_cast<const pl_tile_id *>(dev);
    g.tile_w = sc->elev.tile_w;
    g.face = sc->face;
    g.n_amp = sc->n_amp;
    g.sphere = sc->norm.sphere;
    g.resid_tile_w = resid ? resid->tile_w : 0;
    g.n = n;
    g.root_quad_size = sc->root_quad_size;
    memcpy(g.noise_amp, sc->noise_amp, sizeof(g.noise_amp));
    pl_timing_begin(ctx, PL_K_GENREQ, n);
    gen_requests_ids_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(g);
    pl_timing_end(ctx);
    PL_CUDA(cudaGetLastError());
    ctx->launches += 1;
    /* the smoothstep factor only grows with the level: the shallowest tile of the batch decides for all */
    if (pl_pair_supported(ctx, &sc->elev, &sc->norm, elev, norm)) {
        if ((rc = pl_check_pair_pools(ctx, &sc->elev, &sc->norm, elev, norm, resid, n)) != PL_OK) return rc;
        return pl_launch_pair(ctx, &sc->elev, &sc->norm, elev, norm, resid, n, ctx->gen_ereq, ctx->gen_nreq,
                              norm_level_all_reg(sc->norm.sphere, sc->root_quad_size, min_level));
    }
    return pl_pair_batch_dev(ctx, &sc->elev, &sc->norm, elev, norm, resid, n, ctx->gen_ereq, ctx->gen_nreq);
}

/* ---- several levels of a subtree in ONE launch (pl_produce_levels) ------------------------------------------ */
namespace {

constexpr int kMaxLevelRanges = 24;
struct GenLevelsArgs {
    GenArgs base;                       /* scene, buffers */
    int nranges;
    int first[kMaxLevelRanges + 1];     /* first request of every range */
    int level[kMaxLevelRanges];
    unsigned long long morton0[kMaxLevelRanges], parent_morton0[kMaxLevelRanges];
    int out_slot0[kMaxLevelRanges], parent_slot0[kMaxLevelRanges];
};

__global__ void __launch_bounds__(128) gen_requests_levels_kernel(const GenLevelsArgs g)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.first[g.nranges]) return;
    int r = 0;
    while (i >= g.first[r + 1]) ++r;
    GenArgs a = g.base;
    a.level = g.level[r];
    a.morton0 = g.morton0[r];
    a.parent_morton0 = g.parent_morton0[r];
    a.out_slot0 = g.out_slot0[r];
    a.parent_slot0 = g.parent_slot0[r];
    a.ereq = g.base.ereq + g.first[r];
    a.nreq = g.base.nreq + g.first[r];
    gen_one(a, i - g.first[r]);
    /* ranges after the first wait for their parents, which are tiles of this launch */
    if (r > 0) a.ereq[i - g.first[r]].pad_[0] = 1;
}

}  // namespace

extern "C" int pl_produce_levels(pl_ctx *ctx, const pl_sweep_scene *sc, pl_pool *elev, pl_pool *norm, int nranges,
                                 const pl_level_range *ranges)
{
    if (!ctx || !ranges || !norm) return pl_set_error(PL_ERR_ARG, "NULL argument");
    if (nranges < 1 || nranges > kMaxLevelRanges) return pl_set_error(PL_ERR_ARG, "1 .. %d level ranges per call", kMaxLevelRanges);
    GenLevelsArgs g;
    long long total = 0;
    for (int r = 0; r < nranges; ++r) {
        const pl_level_range &q = ranges[r];
        int rc = check_range(sc, elev, norm, q.level, q.morton0, q.n, q.out_slot0, q.parent_slot0, q.parent_morton0);
        if (rc) return rc;
        if (q.n < 1) return pl_set_error(PL_ERR_ARG, "range %d is empty", r);
        if (r > 0) {
            /* the contract that makes the launch deadlock-free: the parents of range r are tiles of range r - 1 */
            const pl_level_range &p = ranges[r - 1];
            const uint64_t p_first = q.morton0 >> 2, p_last = (q.morton0 + q.n - 1) >> 2;
            if (q.level != p.level + 1 || p_first < p.morton0 || p_last >= p.morton0 + (uint64_t) p.n ||
                q.parent_morton0 != p.morton0 || q.parent_slot0 != p.out_slot0)
                return pl_set_error(PL_ERR_ARG, "range %d: its parents are not the tiles of range %d", r, r - 1);
        }
        g.first[r] = (int) total;
        g.level[r] = q.level;
        g.morton0[r] = q.morton0;
        g.parent_morton0[r] = q.parent_morton0;
        g.out_slot0[r] = q.out_slot0;
        g.parent_slot0[r] = q.parent_slot0;
        total += q.n;
        if (total > (1ll << 30)) return pl_set_error(PL_ERR_ARG, "too many tiles");
    }
    g.first[nranges] = (int) total;
    g.nranges = nranges;
    /* output slots of different ranges must not overlap (each is checked against its own parents by check_range) */
    for (int r = 0; r < nranges; ++r)
        for (int q = 0; q < r; ++q)
            if (ranges[r].out_slot0 < ranges[q].out_slot0 + ranges[q].n && ranges[q].out_slot0 < ranges[r].out_slot0 + ranges[r].n)
                return pl_set_error(PL_ERR_ARG, "ranges %d and %d share output slots", q, r);
    /* RGBA8 normals read the PARENT's normal tile, which a parent of the same launch publishes too late (the ready flag goes
     * up behind its elevation planes): level by level only */
    if (!pl_pair_supported(ctx, &sc->elev, &sc->norm, elev, norm) || norm->kind != PL_POOL_NORM_UN8x2)
        return pl_set_error(PL_ERR_ARG, "pl_produce_levels needs the shipped geometry (101 / 97, grid 4, RG8): use pl_produce_range");
    const int n = (int) total;
    PL_CUDA(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ensure_perlin(ctx)) != PL_OK) return rc;
    if ((rc = ensure_gen_buffers(ctx, n)) != PL_OK) return rc;
    g.base.perlin.perm = ctx->perlin_perm;
    g.base.perlin.g2 = ctx->perlin_g2;
    g.base.ereq = ctx->gen_ereq;
    g.base.nreq = ctx->gen_nreq;
    fill_gen_args(g.base, sc, 0, 0, n, 0, 0, 0);
    pl_timing_begin(ctx, PL_K_GENREQ, n);
    gen_requests_levels_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(g);
    pl_timing_end(ctx);
    PL_CUDA(cudaGetLastError());
    ctx->launches += 1;
    ctx->levels_epoch = ctx->levels_epoch == 0x7fffffff ? 1 : ctx->levels_epoch + 1;
    bool all_reg = true;
    for (int r = 0; r < nranges; ++r) all_reg = all_reg && norm_level_all_reg(sc->norm.sphere, sc->root_quad_size, ranges[r].level);
    return pl_launch_pair_levels(ctx, &sc->elev, &sc->norm, elev, norm, n, ctx->gen_ereq, ctx->gen_nreq, ctx->levels_epoch, all_reg);
}

extern "C" int pl_make_tile_ids_range(int level, uint64_t morton0, int n, int out_slot0, int parent_slot0,
                                      uint64_t parent_morton0, pl_tile_id *ids)
{
    if (!ids || n < 0 || level < 0 || level > 24) return pl_set_error(PL_ERR_ARG, "bad argument");
    for (int i = 0; i < n; ++i) {
        const unsigned long long m = morton0 + (unsigned long long) i;
        pl_tile_id t;
        morton_decode(m, &t.tx, &t.ty);
        t.level = level;
        t.elev_slot = t.norm_slot = out_slot0 + i;
        t.parent_slot = level > 0 ? parent_slot0 + (int) ((m >> 2) - parent_morton0) : -1;
        t.resid_slot = -1;
        t.pad_ = 0;
        ids[i] = t;
    }
    return PL_OK;
}

/* the same requests built on the host (all hardware threads): the per-tile
 * plugin path at batch granularity, and the checker of the device generator */
extern "C" int pl_make_requests_range(const pl_sweep_scene *sc, int level, uint64_t morton0, int n, int out_slot0,
                                      int parent_slot0, uint64_t parent_morton0, pl_elev_req *elev_reqs,
                                      pl_norm_req *norm_reqs, int nthreads)
{
    if (!sc || !elev_reqs || n < 0 || level < 0 || level > 24) return pl_set_error(PL_ERR_ARG, "bad argument");
    GenArgs g;
    g.perlin = pl_host_perlin();
    g.ereq = elev_reqs;
    g.nreq = norm_reqs;
    fill_gen_args(g, sc, level, morton0, n, out_slot0, parent_slot0, parent_morton0);
    if (nthreads <= 0) nthreads = (int) std::thread::hardware_concurrency();
    if (nthreads < 1) nthreads = 1;
    if (n < 4096 || nthreads == 1) {
        for (int i = 0; i < n; ++i) gen_one(g, i);
        return PL_OK;
    }
    std::vector<std::thread> pool;
    const int chunk = (n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; ++t) {
        const int lo = t * chunk, hi = lo + chunk < n ? lo + chunk : n;
        if (lo >= hi) break;
        pool.emplace_back([&g, lo, hi]() { for (int i = lo; i < hi; ++i) gen_one(g, i); });
    }
    for (auto &th : pool) th.join();
    return PL_OK;
}

/* ---- the same for ortho tiles: requests generated on the device, then the ortho kernel ---------------- */

namespace {

struct GenOrthoArgs {
    PerlinView perlin;
    pl_ortho_req *req;
    pl_ortho_scene sc;
    int level, n;
    unsigned long long morton0, parent_morton0;
    int out_slot0, parent_slot0;
};

__global__ void __launch_bounds__(128) gen_ortho_requests_kernel(const GenOrthoArgs g)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    const unsigned long long m = g.morton0 + (unsigned long long) i;
    int tx, ty;
    morton_decode(m, &tx, &ty);
    pl_ortho_req q;
    ortho_fill_req(g.perlin, &g.sc, g.level, tx, ty, &q);
    q.out_slot = g.out_slot0 + i;
    q.parent_slot = g.level > 0 ? g.parent_slot0 + (int) ((m >> 2) - g.parent_morton0) : -1;
    g.req[i] = q;
}

}  // namespace

extern "C" int pl_ortho_produce_range(pl_ctx *ctx, const pl_ortho_scene *sc, pl_pool *ortho, int level, uint64_t morton0,
                                      int n, int out_slot0, int parent_slot0, uint64_t parent_morton0)
{
    if (!ctx || !sc || !ortho) return pl_set_error(PL_ERR_ARG, "NULL argument");
    if (level < 0 || level > 24) return pl_set_error(PL_ERR_ARG, "level %d out of range", level);
    if (n < 0 || morton0 + (uint64_t) n > (1ull << (2 * level))) return pl_set_error(PL_ERR_ARG, "Morton range exceeds level %d", level);
    if (out_slot0 < 0 || out_slot0 + n > ortho->capacity)
        return pl_set_error(PL_ERR_POOL_FULL, "slots [%d,%d) exceed the pool capacity", out_slot0, out_slot0 + n);
    if (sc->n_amp < 0 || sc->n_amp > 32) return pl_set_error(PL_ERR_ARG, "n_amp out of range");
    if (level > 0 && n > 0) {
        const uint64_t p_first = morton0 >> 2, p_last = (morton0 + n - 1) >> 2;
        if (p_first < parent_morton0) return pl_set_error(PL_ERR_ARG, "parent range starts after the first parent");
        const int64_t s_first = (int64_t) parent_slot0 + (int64_t) (p_first - parent_morton0);
        const int64_t s_last = (int64_t) parent_slot0 + (int64_t) (p_last - parent_morton0);
        if (s_first < 0 || s_last >= ortho->capacity) return pl_set_error(PL_ERR_ARG, "parent slots out of range");
        if (s_first < (int64_t) out_slot0 + n && s_last >= out_slot0)
            return pl_set_error(PL_ERR_ARG, "parent slots overlap the output slots");
    }
    if (n == 0) return PL_OK;
    PL_CUDA(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ensure_perlin(ctx)) != PL_OK) return rc;
    if ((rc = ensure_gen_buffers(ctx, n)) != PL_OK) return rc;
    static_assert(sizeof(pl_ortho_req) == sizeof(pl_elev_req), "the generated ortho requests reuse the elevation request buffer");
    GenOrthoArgs g;
    g.perlin.perm = ctx->perlin_perm;
    g.perlin.g2 = ctx->perlin_g2;
    g.req = reinterpret_cast<pl_ortho_req *>(ctx->gen_ereq);
    g.sc = *sc;
    g.level = level;
    g.n = n;
    g.morton0 = morton0;
    g.parent_morton0 = parent_morton0;
    g.out_slot0 = out_slot0;
    g.parent_slot0 = parent_slot0;
    pl_timing_begin(ctx, PL_K_GENREQ, n);
    gen_ortho_requests_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(g);
    pl_timing_end(ctx);
    PL_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return pl_ortho_batch_dev(ctx, sc, ortho, nullptr, n, g.req);
}

extern "C" int pl_debug_download_requests(pl_ctx *ctx, int n, pl_elev_req *elev_reqs, pl_norm_req *norm_reqs)
{
    if (!ctx || n < 0 || n > ctx->gen_cap) return pl_set_error(PL_ERR_ARG, "bad argument");
    PL_CUDA(cudaSetDevice(ctx->device));
    if (elev_reqs)
        PL_CUDA(cudaMemcpyAsync(elev_reqs, ctx->gen_ereq, sizeof(pl_elev_req) * (size_t) n, cudaMemcpyDeviceToHost, ctx->stream));
    if (norm_reqs)
        PL_CUDA(cudaMemcpyAsync(norm_reqs, ctx->gen_nreq, sizeof(pl_norm_req) * (size_t) n, cudaMemcpyDeviceToHost, ctx->stream));
    PL_CUDA(cudaStreamSynchronize(ctx->stream));
    return PL_OK;
}
