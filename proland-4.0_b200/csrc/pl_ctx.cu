/*
 * pl_ctx.cu -- context, device tile pools, noise texture, request staging: the
 * host half of the C ABI in include/proland_b200.h.
 *
 * Replaces producer/GPUTileStorage.cpp:129-199 (a Texture2DArray per pool, a
 * slot = one layer) by one cudaMalloc slab per pool, and the R16F noise
 * Texture2DArray of ElevationProducer.cpp:129-130 by 24 fp16 planes (6 layers x
 * 4 rotations) so that every kernel read of the noise is a coalesced row read.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <thread>
#include <vector>

#include "pl_internal.h"

static thread_local char g_err[512] = "";

int pl_set_error(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

extern "C" const char *pl_last_error(void) { return g_err; }
extern "C" int pl_abi_version(void) { return PL_ABI_VERSION; }

/* ------------------------------------------------------------------ context */

extern "C" int pl_ctx_create(int device, pl_ctx **out)
{
    if (!out) return pl_set_error(PL_ERR_ARG, "pl_ctx_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return pl_set_error(PL_ERR_NO_DEVICE, "no CUDA device (%s); this library has no CPU fallback",
                            e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
    if (device < 0 || device >= count) return pl_set_error(PL_ERR_ARG, "device %d out of range", device);
    PL_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    PL_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return pl_set_error(PL_ERR_NO_DEVICE, "device %d is sm_%d%d; kernels are built for sm_100a only",
                            device, prop.major, prop.minor);
    pl_ctx *ctx = new pl_ctx();
    memset((void *) ctx, 0, sizeof(*ctx));
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    PL_CUDA(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    ctx->stream = ctx->own_stream;
    ctx->timed = new std::vector<pl_ctx::TimedLaunch>();
    ctx->event_pool = new std::vector<cudaEvent_t>();
    *out = ctx;
    return PL_OK;
}

extern "C" void pl_ctx_destroy(pl_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    for (int i = 0; i < ctx->n_noise; ++i) cudaFree(ctx->noise_tabs[i].rot);
    if (ctx->ortho_noise_rot) cudaFree(ctx->ortho_noise_rot);
    for (auto &rb : ctx->readback) {
        if (rb.pinned) cudaFreeHost(rb.pinned);
        if (rb.done) cudaEventDestroy(rb.done);
    }
    if (ctx->stage_fifo) {
        for (auto &e : *ctx->stage_fifo) {
            cudaEventDestroy(e.copied);
            cudaEventDestroy(e.consumed);
        }
        delete ctx->stage_fifo;
    }
    if (ctx->stage_events) {
        for (auto &e : *ctx->stage_events) cudaEventDestroy(e);
        delete ctx->stage_events;
    }
    if (ctx->stage_dev) cudaFree(ctx->stage_dev);
    if (ctx->stage_pinned) cudaFreeHost(ctx->stage_pinned);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->perlin_perm) cudaFree(ctx->perlin_perm);
    if (ctx->perlin_g2) cudaFree(ctx->perlin_g2);
    if (ctx->gen_ereq) cudaFree(ctx->gen_ereq);
    if (ctx->gen_nreq) cudaFree(ctx->gen_nreq);
    if (ctx->resid_scratch) cudaFree(ctx->resid_scratch);
    if (ctx->zgather) cudaFree(ctx->zgather);
    for (auto &t : *ctx->timed) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
    for (auto &e : *ctx->event_pool) cudaEventDestroy(e);
    delete ctx->timed;
    delete ctx->event_pool;
    cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

extern "C" int pl_ctx_set_stream(pl_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return pl_set_error(PL_ERR_ARG, "ctx is NULL");
    ctx->stream = cuda_stream ? (cudaStream_t) cuda_stream : ctx->own_stream;
    return PL_OK;
}

extern "C" void *pl_ctx_stream(pl_ctx *ctx) { return ctx ? (void *) ctx->stream : nullptr; }

extern "C" int pl_sync(pl_ctx *ctx)
{
    if (!ctx) return pl_set_error(PL_ERR_ARG, "ctx is NULL");
    PL_CUDA(cudaStreamSynchronize(ctx->stream));
    return PL_OK;
}

extern "C" uint64_t pl_ctx_launch_count(const pl_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" int pl_device_sm_count(pl_ctx *ctx) { return ctx ? ctx->sm_count : 0; }

/* ------------------------------------------------------------------- timing */

static cudaEvent_t take_event(pl_ctx *ctx)
{
    if (!ctx->event_pool->empty()) {
        cudaEvent_t e = ctx->event_pool->back();
        ctx->event_pool->pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

void pl_timing_begin(pl_ctx *ctx, int kernel, int tiles)
{
    if (!ctx->timing) return;
    pl_ctx::TimedLaunch t;
    t.a = take_event(ctx);
    t.b = take_event(ctx);
    t.kernel = kernel;
    t.tiles = tiles;
    cudaEventRecord(t.a, ctx->stream);
    ctx->timed->push_back(t);
}

void pl_timing_end(pl_ctx *ctx)
{
    if (!ctx->timing) return;
    cudaEventRecord(ctx->timed->back().b, ctx->stream);
}

extern "C" int pl_debug_force_generic(pl_ctx *ctx, int on)
{
    if (!ctx) return pl_set_error(PL_ERR_ARG, "ctx is NULL");
    ctx->force_generic = on ? 1 : 0;
    return PL_OK;
}

extern "C" int pl_debug_stage_ring(pl_ctx *ctx, size_t min_bytes)
{
    if (!ctx) return pl_set_error(PL_ERR_ARG, "ctx is NULL");
    ctx->stage_min = min_bytes;
    return PL_OK;
}

extern "C" int pl_debug_no_fuse(pl_ctx *ctx, int on)
{
    if (!ctx) return pl_set_error(PL_ERR_ARG, "ctx is NULL");
    ctx->no_fuse = on ? 1 : 0;
    return PL_OK;
}

extern "C" int pl_debug_no_slim(pl_ctx *ctx, int on)
{
    if (!ctx) return pl_set_error(PL_ERR_ARG, "ctx is NULL");
    ctx->no_slim = on > 0 ? 1 : 0;
    ctx->slim_sphere = on < 0 ? 1 : 0;
    return PL_OK;
}

extern "C" int pl_timing_enable(pl_ctx *ctx, int on)
{
    if (!ctx) return pl_set_error(PL_ERR_ARG, "ctx is NULL");
    ctx->timing = on ? 1 : 0;
    return PL_OK;
}

extern "C" int pl_timing_collect(pl_ctx *ctx, double *ms, uint64_t *launches, uint64_t *tiles)
{
    if (!ctx || !ms || !launches || !tiles) return pl_set_error(PL_ERR_ARG, "NULL argument");
    PL_CUDA(cudaSetDevice(ctx->device));
    PL_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < PL_K_COUNT; ++k) { ms[k] = 0.0; launches[k] = 0; tiles[k] = 0; }
    for (auto &t : *ctx->timed) {
        float dt = 0.0f;
        if (cudaEventElapsedTime(&dt, t.a, t.b) == cudaSuccess) {
            ms[t.kernel] += dt;
            launches[t.kernel] += 1;
            tiles[t.kernel] += (uint64_t) t.tiles;
        }
        ctx->event_pool->push_back(t.a);
        ctx->event_pool->push_back(t.b);
    }
    ctx->timed->clear();
    return PL_OK;
}

namespace {

struct StageTicket { size_t off, aoff, bytes; };

cudaEvent_t stage_event(pl_ctx *ctx)
{
    if (!ctx->stage_events) ctx->stage_events = new std::vector<cudaEvent_t>();
    if (!ctx->stage_events->empty()) {
        cudaEvent_t e = ctx->stage_events->back();
        ctx->stage_events->pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    return e;
}

/* the oldest entry leaves the ring: its pinned bytes are free when its upload is done, its device bytes
 * when its consumer kernel is done (later uploads are ordered behind that kernel) */
int stage_retire_front(pl_ctx *ctx)
{
    pl_ctx::StageEntry &e = ctx->stage_fifo->front();
    if (!e.consumed_rec) {   /* the newest entry: its consumer has been launched by now (see stage_acquire) */
        PL_CUDA(cudaEventRecord(e.consumed, e.stream));
        e.consumed_rec = 1;
    }
    PL_CUDA(cudaEventSynchronize(e.copied));
    PL_CUDA(cudaStreamWaitEvent(ctx->copy_stream, e.consumed, 0));
    ctx->stage_events->push_back(e.copied);
    ctx->stage_events->push_back(e.consumed);
    ctx->stage_fifo->pop_front();
    return PL_OK;
}

/* room for both arrays in the staging ring; returns where to write them on the host */
int stage_acquire(pl_ctx *ctx, size_t abytes, size_t bbytes, StageTicket *tk, void **apinned, void **bpinned)
{
    PL_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->copy_stream) PL_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    if (!ctx->stage_fifo) ctx->stage_fifo = new std::deque<pl_ctx::StageEntry>();
    /* the kernel that consumes the previous commit has been launched by now (callers launch right after
     * the commit): everything on the stream up to here reads that entry's device bytes */
    if (!ctx->stage_fifo->empty() && !ctx->stage_fifo->back().consumed_rec) {
        PL_CUDA(cudaEventRecord(ctx->stage_fifo->back().consumed, ctx->stage_fifo->back().stream));   /* the stream it was committed for */
        ctx->stage_fifo->back().consumed_rec = 1;
    }
    const size_t aoff = (abytes + 255) & ~(size_t) 255;
    const size_t bytes = (aoff + bbytes + 255) & ~(size_t) 255;
    if (bytes * 3 > ctx->stage_size) {
        /* grow: four of the largest batch seen, at least 32 MB.  Everything in flight must drain first */
        PL_CUDA(cudaStreamSynchronize(ctx->stream));
        PL_CUDA(cudaStreamSynchronize(ctx->copy_stream));
        while (!ctx->stage_fifo->empty()) {
            ctx->stage_events->push_back(ctx->stage_fifo->front().copied);
            ctx->stage_events->push_back(ctx->stage_fifo->front().consumed);
            ctx->stage_fifo->pop_front();
        }
        if (ctx->stage_dev) cudaFree(ctx->stage_dev);
        if (ctx->stage_pinned) cudaFreeHost(ctx->stage_pinned);
        ctx->stage_dev = ctx->stage_pinned = nullptr;
        ctx->stage_size = ctx->stage_w = 0;
        size_t size = bytes * 4;
        const size_t floor_size = ctx->stage_min ? ctx->stage_min : ((size_t) 32 << 20);
        if (size < floor_size) size = floor_size;
        PL_CUDA(cudaMalloc(&ctx->stage_dev, size));
        PL_CUDA(cudaMallocHost(&ctx->stage_pinned, size));
        ctx->stage_size = size;
    }
    /* first fit at the write offset, wrapping once; retire the oldest entries that are in the way */
    size_t off;
    for (;;) {
        std::deque<pl_ctx::StageEntry> &q = *ctx->stage_fifo;
        if (q.empty()) { ctx->stage_w = 0; off = 0; break; }
        const size_t f = q.front().off, w = ctx->stage_w;
        if (f < w || (f == w && false)) {                 /* occupied [f, w): free [w, size) and [0, f) */
            if (w + bytes <= ctx->stage_size) { off = w; break; }
            if (bytes <= f) { off = 0; break; }
        } else {                                          /* occupied [f, size) + [0, w): free [w, f) */
            if (w + bytes <= f) { off = w; break; }
        }
        int rc = stage_retire_front(ctx);
        if (rc) return rc;
    }
    tk->off = off;
    tk->aoff = aoff;
    tk->bytes = bytes;
    *apinned = static_cast<char *>(ctx->stage_pinned) + off;
    if (bpinned) *bpinned = static_cast<char *>(ctx->stage_pinned) + off + aoff;
    return PL_OK;
}

/* enqueue the upload of a filled range on the copy stream -- beside the kernels of earlier batches --
 * and make the kernel stream wait for it; returns the device addresses of the two arrays */
int stage_commit(pl_ctx *ctx, const StageTicket &tk, void **adev, void **bdev)
{
    pl_ctx::StageEntry e;
    e.off = tk.off;
    e.bytes = tk.bytes;
    e.copied = stage_event(ctx);
    e.consumed = stage_event(ctx);
    e.consumed_rec = 0;
    e.stream = ctx->stream;
    if (!e.copied || !e.consumed) return pl_set_error(PL_ERR_CUDA, "cudaEventCreate failed");
    char *d = static_cast<char *>(ctx->stage_dev) + tk.off;
    PL_CUDA(cudaMemcpyAsync(d, static_cast<char *>(ctx->stage_pinned) + tk.off, tk.bytes, cudaMemcpyHostToDevice,
                            ctx->copy_stream));
    PL_CUDA(cudaEventRecord(e.copied, ctx->copy_stream));
    PL_CUDA(cudaStreamWaitEvent(ctx->stream, e.copied, 0));
    ctx->stage_fifo->push_back(e);
    ctx->stage_w = tk.off + tk.bytes;
    *adev = d;
    if (bdev) *bdev = d + tk.aoff;
    return PL_OK;
}

/* f(lo, hi) over [0, n) on up to 16 host threads (the calling thread takes the first chunk) */
template <class F>
void parallel_chunks(int n, int min_chunk, F f)
{
    int nt = (int) std::thread::hardware_concurrency();
    nt = nt < 1 ? 1 : (nt > 16 ? 16 : nt);
    /* several ranks share one box: PL_HOST_THREADS caps the threads of one process */
    static const int cap = []() { const char *e = getenv("PL_HOST_THREADS"); return e ? atoi(e) : 0; }();
    if (cap > 0 && nt > cap) nt = cap;
    if (nt > n / min_chunk) nt = n / min_chunk;
    if (nt <= 1) { f(0, n); return; }
    const int chunk = (n + nt - 1) / nt;
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) {
        const int lo = t * chunk, hi = lo + chunk < n ? lo + chunk : n;
        if (lo < hi) pool.emplace_back([=]() { f(lo, hi); });
    }
    f(0, chunk < n ? chunk : n);
    for (auto &th : pool) th.join();
}

}  // namespace

int pl_stage_requests2(pl_ctx *ctx, const void *a, size_t abytes, const void *b, size_t bbytes, void **adev, void **bdev)
{
    StageTicket tk;
    void *pa = nullptr, *pb = nullptr;
    int rc = stage_acquire(ctx, abytes, b ? bbytes : 0, &tk, &pa, &pb);
    if (rc) return rc;
    memcpy(pa, a, abytes);
    if (b) memcpy(pb, b, bbytes);
    return stage_commit(ctx, tk, adev, bdev);
}

/* -------------------------------------------------------------------- pools */

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_tiled()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn) p;
    }
    return fn;
}

static int elem_bytes(int kind)
{
    switch (kind) {
    case PL_POOL_ELEV_F32x3: return 12;
    case PL_POOL_NORM_UN8x2: return 2;
    case PL_POOL_NORM_UN8x4: return 4;
    case PL_POOL_RESID_F32: return 4;
    case PL_POOL_RESID_I16: return 2;
    case PL_POOL_ORTHO_UN8x4: return 4;
    default: return 0;
    }
}

static thread_local bool g_pool_shared = false;

extern "C" int pl_pool_create_shared(pl_ctx *ctx, int kind, int tile_w, int capacity, pl_pool **out)
{
    g_pool_shared = true;
    const int rc = pl_pool_create(ctx, kind, tile_w, capacity, out);
    g_pool_shared = false;
    return rc;
}

extern "C" int pl_pool_create(pl_ctx *ctx, int kind, int tile_w, int capacity, pl_pool **out)
{
    if (!ctx || !out) return pl_set_error(PL_ERR_ARG, "pl_pool_create: NULL argument");
    *out = nullptr;
    if (elem_bytes(kind) == 0) return pl_set_error(PL_ERR_ARG, "unknown pool kind %d", kind);
    if (tile_w < 8 || tile_w > 1024 || capacity <= 0)
        return pl_set_error(PL_ERR_ARG, "bad pool geometry tile_w=%d capacity=%d", tile_w, capacity);
    PL_CUDA(cudaSetDevice(ctx->device));
    pl_pool *p = new pl_pool();
    memset(p, 0, sizeof(*p));
    p->ctx = ctx;
    p->kind = kind;
    p->tile_w = tile_w;
    p->capacity = capacity;
    p->tile_bytes = (size_t) tile_w * tile_w * elem_bytes(kind);
    switch (kind) {
    case PL_POOL_ELEV_F32x3:
        if ((tile_w - 5) % 2 != 0 || tile_w < 9) {
            delete p;
            return pl_set_error(PL_ERR_ARG, "elevation tile_w must be odd (tileSize + 5), got %d", tile_w);
        }
        p->pitch = pl_round_up(tile_w, 4);
        p->plane_elems = (size_t) tile_w * p->pitch;
        p->slot_bytes = 3 * p->plane_elems * sizeof(float);
        break;
    case PL_POOL_NORM_UN8x2:
    case PL_POOL_NORM_UN8x4:
        p->pitch = tile_w * elem_bytes(kind);
        p->slot_bytes = (size_t) pl_round_up((int) (((size_t) tile_w * p->pitch + 15) / 16 * 16), 32);
        break;
    case PL_POOL_RESID_F32:
        p->pitch = pl_round_up(tile_w, 4);
        p->slot_bytes = (size_t) pl_round_up(tile_w * p->pitch * 4, 128);
        break;
    case PL_POOL_RESID_I16:
        p->pitch = pl_round_up(tile_w, 8);
        p->slot_bytes = (size_t) pl_round_up(tile_w * p->pitch * 2, 128);
        break;
    case PL_POOL_ORTHO_UN8x4:
        if ((tile_w - 4) % 8 != 0) {
            delete p;
            return pl_set_error(PL_ERR_ARG, "ortho tile_w - 4 must be a multiple of 8 (16-byte quadrant rows), got %d", tile_w);
        }
        p->pitch = tile_w * 4;
        p->slot_bytes = (size_t) pl_round_up(tile_w * p->pitch, 128);
        break;
    }
    /* F32 residual pools carry one hidden scratch slot (index capacity, PL_SLOT_SCRATCH): the
     * temporary of the root-level composition, ResidualProducer.cpp:218-228 */
    const size_t nslots = (size_t) capacity + (kind == PL_POOL_RESID_F32 ? 1 : 0);
    cudaError_t e = cudaSuccess;
    if (g_pool_shared) {
        /* pl_pool_create_shared: the memory comes from the VMM allocator with a shareable handle (pl_multicast.cu) */
        const int rc = pl_vmm_alloc(p, p->slot_bytes * nslots);
        if (rc) {
            delete p;
            return rc;
        }
    } else {
        e = cudaMalloc(&p->base, p->slot_bytes * nslots);
    }
    if (e != cudaSuccess) {
        const size_t want = p->slot_bytes * nslots;
        delete p;
        cudaGetLastError();
        return pl_set_error(PL_ERR_POOL_FULL, "cudaMalloc of %zu bytes for %d slots failed: %s", want, capacity,
                            cudaGetErrorString(e));
    }
    /* pad columns / bytes are part of whole-sector writes later; start from zero */
    PL_CUDA(cudaMemsetAsync(p->base, 0, p->slot_bytes * nslots, ctx->stream));

    if (kind == PL_POOL_ELEV_F32x3) {
        PL_CUDA(cudaMalloc(&p->stats, sizeof(float2) * capacity));
        PL_CUDA(cudaMalloc(&p->ready, sizeof(int) * capacity));
        PL_CUDA(cudaMemset(p->ready, 0, sizeof(int) * capacity));
        PL_CUDA(cudaMemsetAsync(p->stats, 0, sizeof(float2) * capacity, ctx->stream));
        /* parent window staged by TMA: (tileSize/2 + 6) texels each way, the
         * inner extent rounded up to 4 floats (16-byte box rows) */
        p->box_h = (tile_w - 5) / 2 + 6;
        p->box_w = pl_round_up(p->box_h, 4);
        EncodeTiledFn enc = get_encode_tiled();
        if (!enc) {
            pl_pool_destroy(p);
            return pl_set_error(PL_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
        }
        /* x: tile_w texels (pitch-padded rows), y: tile_w rows, z: plane index
         * (3 planes per slot, slots are contiguous) */
        cuuint64_t dims[3] = { (cuuint64_t) tile_w, (cuuint64_t) tile_w, (cuuint64_t) capacity * 3 };
        cuuint64_t strides[2] = { (cuuint64_t) p->pitch * 4, (cuuint64_t) p->plane_elems * 4 };
        cuuint32_t box[3] = { (cuuint32_t) p->box_w, (cuuint32_t) p->box_h, 1 };
        cuuint32_t estr[3] = { 1, 1, 1 };
        CUresult r = enc(&p->tm_parent, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, p->base, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            pl_pool_destroy(p);
            return pl_set_error(PL_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int) r);
        }
    }
    *out = p;
    return PL_OK;
}

extern "C" void pl_pool_destroy(pl_pool *p)
{
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    cudaStreamSynchronize(p->ctx->stream);
    for (int i = 0; i < p->npeers; ++i) cudaIpcCloseMemHandle(p->peer_base[i]);
    if (p->vmm) pl_vmm_free(p);
    else if (p->base) cudaFree(p->base);
    if (p->stats) cudaFree(p->stats);
    if (p->ready) cudaFree(p->ready);
    delete p;
}

/* ---- peers: the copies of a pool on the other GPUs of the box ---------------------------------- */

extern "C" int pl_pool_export(pl_pool *p, void *handle64)
{
    if (!p || !handle64) return pl_set_error(PL_ERR_ARG, "NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == PL_IPC_HANDLE_BYTES, "handle size");
    PL_CUDA(cudaSetDevice(p->ctx->device));
    cudaIpcMemHandle_t h;
    if (p->vmm) return pl_set_error(PL_ERR_ARG, "a shared (VMM) pool has no CUDA IPC handle: use the multicast entry points");
    PL_CUDA(cudaIpcGetMemHandle(&h, p->base));
    memcpy(handle64, &h, sizeof(h));
    return PL_OK;
}

extern "C" int pl_pool_attach_peers(pl_pool *p, int n, const void *handles, int self)
{
    if (!p || !handles || n < 1 || n > pl_pool::kMaxPeers + 1 || self < 0 || self >= n)
        return pl_set_error(PL_ERR_ARG, "pl_pool_attach_peers: %d handles, self %d (at most %d GPUs)", n, self, pl_pool::kMaxPeers + 1);
    if (p->npeers) return pl_set_error(PL_ERR_ARG, "peers are already attached");
    PL_CUDA(cudaSetDevice(p->ctx->device));
    for (int i = 0; i < n; ++i) {
        if (i == self) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *) handles + (size_t) i * PL_IPC_HANDLE_BYTES, sizeof(h));
        void *ptr = nullptr;
        PL_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        p->peer_base[p->npeers++] = (uint8_t *) ptr;
    }
    return PL_OK;
}

extern "C" int pl_pool_push_to_peers(pl_pool *p, int on)
{
    if (!p) return pl_set_error(PL_ERR_ARG, "pool is NULL");
    if (on && p->kind != PL_POOL_NORM_UN8x2) return pl_set_error(PL_ERR_ARG, "only RG8 normal pools push their tiles");
    if (on == 2 && p->mc_state != 3) return pl_set_error(PL_ERR_ARG, "the pool is not bound to a multicast object (pl_pool_mc_bind)");
    if (on != 2 && on && !p->npeers) return pl_set_error(PL_ERR_ARG, "no peers attached");
    p->push = on == 2 ? 2 : (on ? 1 : 0);
    return PL_OK;
}

extern "C" int pl_pool_capacity(const pl_pool *p) { return p ? p->capacity : 0; }
extern "C" int pl_pool_tile_w(const pl_pool *p) { return p ? p->tile_w : 0; }
extern "C" size_t pl_pool_tile_bytes(const pl_pool *p) { return p ? p->tile_bytes : 0; }
extern "C" size_t pl_pool_slot_bytes(const pl_pool *p) { return p ? p->slot_bytes : 0; }
extern "C" void *pl_pool_device_ptr(pl_pool *p) { return p ? p->base : nullptr; }

static int check_slot(const pl_pool *p, int slot, size_t bytes)
{
    if (!p) return pl_set_error(PL_ERR_ARG, "pool is NULL");
    if (slot < 0 || slot >= p->capacity) return pl_set_error(PL_ERR_ARG, "slot %d out of range [0,%d)", slot, p->capacity);
    if (bytes != p->tile_bytes) return pl_set_error(PL_ERR_ARG, "expected %zu bytes per tile, got %zu", p->tile_bytes, bytes);
    return PL_OK;
}

extern "C" int pl_pool_download(pl_pool *p, int slot, void *host, size_t bytes)
{
    int rc = check_slot(p, slot, bytes);
    if (rc) return rc;
    if (!host) return pl_set_error(PL_ERR_ARG, "host is NULL");
    PL_CUDA(cudaSetDevice(p->ctx->device));
    cudaStream_t s = p->ctx->stream;
    const uint8_t *src = p->base + (size_t) slot * p->slot_bytes;
    const int W = p->tile_w;
    switch (p->kind) {
    case PL_POOL_ELEV_F32x3: {
        std::vector<float> planes(3 * p->plane_elems);
        PL_CUDA(cudaMemcpyAsync(planes.data(), src, p->slot_bytes, cudaMemcpyDeviceToHost, s));
        PL_CUDA(cudaStreamSynchronize(s));
        float *o = (float *) host;
        for (int y = 0; y < W; ++y)
            for (int x = 0; x < W; ++x)
                for (int c = 0; c < 3; ++c)
                    o[(size_t) (x + y * W) * 3 + c] = planes[c * p->plane_elems + (size_t) y * p->pitch + x];
        break;
    }
    case PL_POOL_NORM_UN8x2:
    case PL_POOL_NORM_UN8x4:
    case PL_POOL_ORTHO_UN8x4:
        PL_CUDA(cudaMemcpyAsync(host, src, p->tile_bytes, cudaMemcpyDeviceToHost, s));
        PL_CUDA(cudaStreamSynchronize(s));
        break;
    case PL_POOL_RESID_F32:
    case PL_POOL_RESID_I16: {
        const size_t eb = p->kind == PL_POOL_RESID_F32 ? 4 : 2;
        PL_CUDA(cudaMemcpy2DAsync(host, W * eb, src, p->pitch * eb, W * eb, W, cudaMemcpyDeviceToHost, s));
        PL_CUDA(cudaStreamSynchronize(s));
        break;
    }
    }
    return PL_OK;
}

/* n consecutive slots in ONE device-to-host copy, then the reference layout per tile on the host (threads): the bulk form of
 * pl_pool_download for callers that read whole levels back (the residual builder) */
extern "C" int pl_pool_download_range(pl_pool *p, int slot0, int n, void *host, size_t bytes)
{
    if (!p || n < 0) return pl_set_error(PL_ERR_ARG, "bad argument");
    if (n == 0) return PL_OK;
    if (!host) return pl_set_error(PL_ERR_ARG, "host is NULL");
    if (slot0 < 0 || (long long) slot0 + n > p->capacity) return pl_set_error(PL_ERR_ARG, "slots [%d, %d) out of range", slot0, slot0 + n);
    if (bytes != p->tile_bytes * (size_t) n) return pl_set_error(PL_ERR_ARG, "buffer of %zu bytes, %d tiles of %zu bytes expected", bytes, n, p->tile_bytes);
    PL_CUDA(cudaSetDevice(p->ctx->device));
    std::vector<uint8_t> stage(p->slot_bytes * (size_t) n);
    PL_CUDA(cudaMemcpyAsync(stage.data(), p->base + (size_t) slot0 * p->slot_bytes, stage.size(), cudaMemcpyDeviceToHost, p->ctx->stream));
    PL_CUDA(cudaStreamSynchronize(p->ctx->stream));
    const int W = p->tile_w;
    uint8_t *out = static_cast<uint8_t *>(host);
    parallel_chunks(n, 8, [&](int lo, int hi) {
        for (int t = lo; t < hi; ++t) {
            const uint8_t *src = stage.data() + (size_t) t * p->slot_bytes;
            uint8_t *dst = out + (size_t) t * p->tile_bytes;
            switch (p->kind) {
            case PL_POOL_ELEV_F32x3: {
                const float *planes = reinterpret_cast<const float *>(src);
                float *o = reinterpret_cast<float *>(dst);
                for (int y = 0; y < W; ++y)
                    for (int x = 0; x < W; ++x)
                        for (int c = 0; c < 3; ++c)
                            o[(size_t) (x + y * W) * 3 + c] = planes[c * p->plane_elems + (size_t) y * p->pitch + x];
                break;
            }
            case PL_POOL_RESID_F32:
            case PL_POOL_RESID_I16: {
                const size_t eb = p->kind == PL_POOL_RESID_F32 ? 4 : 2;
                for (int y = 0; y < W; ++y) memcpy(dst + (size_t) y * W * eb, src + (size_t) y * p->pitch * eb, W * eb);
                break;
            }
            default:
                memcpy(dst, src, p->tile_bytes);
                break;
            }
        }
    });
    return PL_OK;
}

extern "C" int pl_pool_upload(pl_pool *p, int slot, const void *host, size_t bytes)
{
    int rc = check_slot(p, slot, bytes);
    if (rc) return rc;
    if (!host) return pl_set_error(PL_ERR_ARG, "host is NULL");
    PL_CUDA(cudaSetDevice(p->ctx->device));
    cudaStream_t s = p->ctx->stream;
    uint8_t *dst = p->base + (size_t) slot * p->slot_bytes;
    const int W = p->tile_w;
    switch (p->kind) {
    case PL_POOL_ELEV_F32x3: {
        std::vector<float> planes(3 * p->plane_elems, 0.0f);
        const float *in = (const float *) host;
        for (int y = 0; y < W; ++y)
            for (int x = 0; x < W; ++x)
                for (int c = 0; c < 3; ++c)
                    planes[c * p->plane_elems + (size_t) y * p->pitch + x] = in[(size_t) (x + y * W) * 3 + c];
        PL_CUDA(cudaMemcpyAsync(dst, planes.data(), p->slot_bytes, cudaMemcpyHostToDevice, s));
        PL_CUDA(cudaStreamSynchronize(s));
        break;
    }
    case PL_POOL_NORM_UN8x2:
    case PL_POOL_NORM_UN8x4:
    case PL_POOL_ORTHO_UN8x4:
        PL_CUDA(cudaMemcpyAsync(dst, host, p->tile_bytes, cudaMemcpyHostToDevice, s));
        PL_CUDA(cudaStreamSynchronize(s));
        break;
    case PL_POOL_RESID_F32:
    case PL_POOL_RESID_I16: {
        const size_t eb = p->kind == PL_POOL_RESID_F32 ? 4 : 2;
        PL_CUDA(cudaMemcpy2DAsync(dst, p->pitch * eb, host, W * eb, W * eb, W, cudaMemcpyHostToDevice, s));
        PL_CUDA(cudaStreamSynchronize(s));
        break;
    }
    }
    return PL_OK;
}

/* -------------------------------------------------------------------- noise */

extern "C" int pl_noise_init(pl_ctx *ctx, int tile_w, float *host_out)
{
    if (!ctx) return pl_set_error(PL_ERR_ARG, "ctx is NULL");
    if (tile_w < 11 || tile_w > 1024) return pl_set_error(PL_ERR_ARG, "bad noise tile_w %d", tile_w);
    PL_CUDA(cudaSetDevice(ctx->device));
    const int W = tile_w;
    const bool have = ctx->noise_for(W) != nullptr;
    if (have && !host_out) return PL_OK;      /* cached per width */
    if (!have && ctx->n_noise == (int) (sizeof(ctx->noise_tabs) / sizeof(ctx->noise_tabs[0])))
        return pl_set_error(PL_ERR_ARG, "too many distinct noise tile widths in one context");
    std::vector<float> n6((size_t) 6 * W * W);
    pl_host_dem_noise(W, n6.data());
    /* the reference uploads the array as R16F: round to nearest even */
    std::vector<__half> h6(n6.size());
    for (size_t i = 0; i < n6.size(); ++i) {
        h6[i] = __float2half_rn(n6[i]);
        if (host_out) host_out[i] = __half2float(h6[i]);
    }
    if (have) return PL_OK;                   /* host_out filled; the device table exists */
    const int pitch = pl_round_up(W + 1, 8);
    std::vector<__half> rot((size_t) 24 * W * pitch, __float2half_rn(0.0f));
    for (int r = 0; r < 4; ++r)
        for (int l = 0; l < 6; ++l)
            for (int y = 0; y < W; ++y)
                for (int x = 0; x < W; ++x) {
                    /* uvs = (nx, ny, 1-nx, 1-ny)[r], [(r+1)%4]  (upsampleShader.glsl:172-174) */
                    int sx, sy;
                    switch (r) {
                    case 0: sx = x; sy = y; break;
                    case 1: sx = y; sy = W - 1 - x; break;
                    case 2: sx = W - 1 - x; sy = W - 1 - y; break;
                    default: sx = W - 1 - y; sy = x; break;
                    }
                    rot[((size_t) (r * 6 + l) * W + y) * pitch + pl_noise_col(x)] = h6[(size_t) l * W * W + sx + (size_t) sy * W];
                }
    __half *dev = nullptr;
    PL_CUDA(cudaMalloc(&dev, rot.size() * sizeof(__half)));
    cudaError_t e = cudaMemcpyAsync(dev, rot.data(), rot.size() * sizeof(__half), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        cudaFree(dev);
        return pl_set_error(PL_ERR_CUDA, "pl_noise_init: %s", cudaGetErrorString(e));
    }
    pl_ctx::NoiseTable &t = ctx->noise_tabs[ctx->n_noise++];
    t.w = W;
    t.pitch = pitch;
    t.rot = dev;
    return PL_OK;
}

/* ----------------------------------------------------------- batch wrappers */

static int check_elev_args(pl_ctx *ctx, const pl_elev_scene *sc, pl_pool *elev, pl_pool *resid, int n)
{
    if (!ctx || !sc || !elev) return pl_set_error(PL_ERR_ARG, "NULL argument");
    if (n < 0) return pl_set_error(PL_ERR_ARG, "n < 0");
    if (elev->kind != PL_POOL_ELEV_F32x3) return pl_set_error(PL_ERR_ARG, "elev is not an ELEV_F32x3 pool");
    if (elev->tile_w != sc->tile_w) return pl_set_error(PL_ERR_ARG, "scene tile_w %d != pool tile_w %d", sc->tile_w, elev->tile_w);
    if (resid && resid->kind != PL_POOL_RESID_F32 && resid->kind != PL_POOL_RESID_I16)
        return pl_set_error(PL_ERR_ARG, "resid is not a residual pool");
    if (resid && (resid->tile_w - 5) % (sc->tile_w - 5) != 0)
        return pl_set_error(PL_ERR_ARG, "residual tile size %d is not a multiple of %d", resid->tile_w - 5, sc->tile_w - 5);
    if (sc->grid <= 0) return pl_set_error(PL_ERR_ARG, "grid must be > 0");
    if (!ctx->noise_for(sc->tile_w))
        return pl_set_error(PL_ERR_ARG, "pl_noise_init(ctx, %d) has not been called", sc->tile_w);
    return PL_OK;
}

extern "C" int pl_elevation_batch_dev(pl_ctx *ctx, const pl_elev_scene *sc, pl_pool *elev, pl_pool *resid, int n,
                                      const pl_elev_req *dev_reqs)
{
    int rc = check_elev_args(ctx, sc, elev, resid, n);
    if (rc) return rc;
    if (n == 0) return PL_OK;
    if (!dev_reqs) return pl_set_error(PL_ERR_ARG, "reqs is NULL");
    return pl_launch_elevation(ctx, sc, elev, resid, n, dev_reqs);
}

/* everything a kernel derives an address from: slots, noise indices, the parent window (dx, dy) and the residual window
 * (rx, ry; the specialised kernel reads it with 16- / 8-byte loads: rx must be a multiple of 4) */
static inline bool elev_req_bad(const pl_elev_req &q, const pl_pool *elev, const pl_pool *resid)
{
    const int rcap = resid ? resid->capacity : 0;
    const int half = (elev->tile_w - 5) / 2;
    if (q.out_slot < 0 || q.out_slot >= elev->capacity || q.parent_slot < -1 || q.parent_slot >= elev->capacity ||
        q.resid_slot < -1 || q.resid_slot >= rcap || q.noise_r < 0 || q.noise_r > 3 || q.noise_l < 0 || q.noise_l > 5 ||
        q.parent_slot == q.out_slot)
        return true;
    if ((q.dx != 0 && q.dx != half) || (q.dy != 0 && q.dy != half)) return true;
    if (q.resid_slot >= 0 &&
        (q.rx < 0 || q.ry < 0 || q.rx + elev->tile_w > resid->tile_w || q.ry + elev->tile_w > resid->tile_w || (q.rx & 3) != 0))
        return true;
    return false;
}

static int check_elev_reqs(const pl_pool *elev, const pl_pool *resid, int n, const pl_elev_req *reqs)
{
    for (int i = 0; i < n; ++i)
        if (elev_req_bad(reqs[i], elev, resid))
            return pl_set_error(PL_ERR_ARG, "request %d: slot, noise index or window out of range", i);
    return PL_OK;
}

extern "C" int pl_elevation_batch(pl_ctx *ctx, const pl_elev_scene *sc, pl_pool *elev, pl_pool *resid, int n,
                                  const pl_elev_req *reqs)
{
    int rc = check_elev_args(ctx, sc, elev, resid, n);
    if (rc) return rc;
    if (n == 0) return PL_OK;
    if (!reqs) return pl_set_error(PL_ERR_ARG, "reqs is NULL");
    if ((rc = check_elev_reqs(elev, resid, n, reqs)) != PL_OK) return rc;
    void *dev = nullptr;
    rc = pl_stage_requests(ctx, reqs, sizeof(pl_elev_req) * (size_t) n, &dev);
    if (rc) return rc;
    return pl_launch_elevation(ctx, sc, elev, resid, n, (const pl_elev_req *) dev);
}

extern "C" int pl_elev_stats_download(pl_ctx *ctx, pl_pool *elev, int n, const int32_t *slots, float *out)
{
    if (!ctx || !elev || !slots || !out || n < 0) return pl_set_error(PL_ERR_ARG, "bad argument");
    if (elev->kind != PL_POOL_ELEV_F32x3) return pl_set_error(PL_ERR_ARG, "not an elevation pool");
    PL_CUDA(cudaSetDevice(ctx->device));
    std::vector<float2> all(elev->capacity);
    PL_CUDA(cudaMemcpyAsync(all.data(), elev->stats, sizeof(float2) * elev->capacity, cudaMemcpyDeviceToHost, ctx->stream));
    PL_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < n; ++i) {
        if (slots[i] < 0 || slots[i] >= elev->capacity) return pl_set_error(PL_ERR_ARG, "slot out of range");
        out[2 * i] = all[slots[i]].x;
        out[2 * i + 1] = all[slots[i]].y;
    }
    return PL_OK;
}

extern "C" int pl_elev_stats_range(pl_ctx *ctx, pl_pool *elev, int slot0, int n, float *out)
{
    if (!ctx || !elev || !out || n < 0 || slot0 < 0) return pl_set_error(PL_ERR_ARG, "bad argument");
    if (elev->kind != PL_POOL_ELEV_F32x3) return pl_set_error(PL_ERR_ARG, "not an elevation pool");
    if (slot0 + n > elev->capacity) return pl_set_error(PL_ERR_ARG, "slot range out of the pool");
    PL_CUDA(cudaSetDevice(ctx->device));
    PL_CUDA(cudaMemcpyAsync(out, elev->stats + slot0, sizeof(float2) * (size_t) n, cudaMemcpyDeviceToHost, ctx->stream));
    PL_CUDA(cudaStreamSynchronize(ctx->stream));
    return PL_OK;
}

/* Asynchronous read-back of per-tile (zmin, zmax): the copy is enqueued behind the kernels already on
 * the stream, the caller collects it later -- the reference reads tile z ranges back through PBOs a few
 * frames late (TileSamplerZ.cpp:253-351, ReadbackManager). */
extern "C" int pl_elev_stats_readback_begin(pl_ctx *ctx, pl_pool *elev, int slot0, int n, int *ticket)
{
    if (!ctx || !elev || !ticket || n < 0 || slot0 < 0) return pl_set_error(PL_ERR_ARG, "bad argument");
    if (elev->kind != PL_POOL_ELEV_F32x3) return pl_set_error(PL_ERR_ARG, "not an elevation pool");
    if (slot0 + n > elev->capacity) return pl_set_error(PL_ERR_ARG, "slot range out of the pool");
    PL_CUDA(cudaSetDevice(ctx->device));
    int k = -1;
    for (int i = 0; i < pl_ctx::kReadbacks; ++i)
        if (!ctx->readback[i].busy) { k = i; break; }
    if (k < 0) return pl_set_error(PL_ERR_ARG, "all %d read-backs are in flight: collect one first", (int) pl_ctx::kReadbacks);
    pl_ctx::Readback &rb = ctx->readback[k];
    if (!rb.done) PL_CUDA(cudaEventCreateWithFlags(&rb.done, cudaEventDisableTiming));
    const size_t bytes = sizeof(float2) * (size_t) n;
    if (bytes > rb.cap) {
        if (rb.pinned) cudaFreeHost(rb.pinned);
        rb.pinned = nullptr;
        rb.cap = 0;
        PL_CUDA(cudaMallocHost(&rb.pinned, bytes + bytes / 2 + 4096));
        rb.cap = bytes + bytes / 2 + 4096;
    }
    if (n) PL_CUDA(cudaMemcpyAsync(rb.pinned, elev->stats + slot0, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    PL_CUDA(cudaEventRecord(rb.done, ctx->stream));
    rb.n = n;
    rb.busy = 1;
    *ticket = k;
    return PL_OK;
}

/* TileSamplerZ's frame read-back (core/.../terrain/TileSamplerZ.cpp:253-351): the z range of up to kZTiles tiles in
 * ARBITRARY slots -- the tiles a frame picked from its needReadback set -- and the zm texel under the camera, gathered by
 * one small kernel and copied back asynchronously like pl_elev_stats_readback_begin.  Result pairs, in this order:
 * [(h, h) of the camera texel, when cam_slot >= 0], then (zmin, zmax) of slots[0..n). */
namespace {
constexpr int kZTiles = 64;
struct ZGather { int n; int cam_slot, cam_x, cam_y; int slots[kZTiles]; };
__global__ void z_gather_kernel(const ZGather g, const float2 *stats, const float *elev, size_t slot_elems, size_t plane, int pitch,
                                float2 *out)
{
    const int i = threadIdx.x;
    const int c = g.cam_slot >= 0 ? 1 : 0;
    if (i == 0 && c) {
        const float h = elev[(size_t) g.cam_slot * slot_elems + 2 * plane + (size_t) g.cam_y * pitch + g.cam_x];
        out[0] = make_float2(h, h);
    }
    if (i < g.n) out[c + i] = stats[g.slots[i]];
}
}  // namespace

extern "C" int pl_elev_zreadback_begin(pl_ctx *ctx, pl_pool *elev, int n, const int32_t *slots, int cam_slot, int cam_x, int cam_y,
                                       int *ticket)
{
    if (!ctx || !elev || !ticket || n < 0 || n > kZTiles || (n > 0 && !slots)) return pl_set_error(PL_ERR_ARG, "bad argument");
    if (elev->kind != PL_POOL_ELEV_F32x3) return pl_set_error(PL_ERR_ARG, "not an elevation pool");
    ZGather g;
    g.n = n;
    g.cam_slot = cam_slot >= 0 ? cam_slot : -1;
    g.cam_x = cam_x;
    g.cam_y = cam_y;
    for (int i = 0; i < n; ++i) {
        if (slots[i] < 0 || slots[i] >= elev->capacity) return pl_set_error(PL_ERR_ARG, "slot out of range");
        g.slots[i] = slots[i];
    }
    if (cam_slot >= elev->capacity || (cam_slot >= 0 && (cam_x < 0 || cam_y < 0 || cam_x >= elev->tile_w || cam_y >= elev->tile_w)))
        return pl_set_error(PL_ERR_ARG, "camera texel out of range");
    PL_CUDA(cudaSetDevice(ctx->device));
    int k = -1;
    for (int i = 0; i < pl_ctx::kReadbacks; ++i)
        if (!ctx->readback[i].busy) { k = i; break; }
    if (k < 0) return pl_set_error(PL_ERR_ARG, "all %d read-backs are in flight: collect one first", (int) pl_ctx::kReadbacks);
    pl_ctx::Readback &rb = ctx->readback[k];
    if (!rb.done) PL_CUDA(cudaEventCreateWithFlags(&rb.done, cudaEventDisableTiming));
    const int pairs = n + (cam_slot >= 0 ? 1 : 0);
    const size_t bytes = sizeof(float2) * (size_t) (kZTiles + 1);
    if (bytes > rb.cap) {
        if (rb.pinned) cudaFreeHost(rb.pinned);
        rb.pinned = nullptr;
        rb.cap = 0;
        PL_CUDA(cudaMallocHost(&rb.pinned, bytes + 4096));
        rb.cap = bytes + 4096;
    }
    if (!ctx->zgather) PL_CUDA(cudaMalloc(&ctx->zgather, sizeof(float2) * (kZTiles + 1) * pl_ctx::kReadbacks));
    float2 *dev = static_cast<float2 *>(ctx->zgather) + (size_t) k * (kZTiles + 1);
    if (pairs) {
        z_gather_kernel<<<1, kZTiles, 0, ctx->stream>>>(g, elev->stats, reinterpret_cast<const float *>(elev->base),
                                                       elev->slot_bytes / sizeof(float), elev->plane_elems, elev->pitch, dev);
        PL_CUDA(cudaGetLastError());
        ctx->launches += 1;
        PL_CUDA(cudaMemcpyAsync(rb.pinned, dev, sizeof(float2) * (size_t) pairs, cudaMemcpyDeviceToHost, ctx->stream));
    }
    PL_CUDA(cudaEventRecord(rb.done, ctx->stream));
    rb.n = pairs;
    rb.busy = 1;
    *ticket = k;
    return PL_OK;
}

/* 1 when the read-back behind `ticket` has arrived (pl_elev_stats_readback_end will not wait), 0 when not, < 0: error */
extern "C" int pl_elev_stats_readback_ready(pl_ctx *ctx, int ticket)
{
    if (!ctx || ticket < 0 || ticket >= pl_ctx::kReadbacks || !ctx->readback[ticket].busy) return -1;
    return cudaEventQuery(ctx->readback[ticket].done) == cudaSuccess ? 1 : 0;
}

extern "C" int pl_elev_stats_readback_end(pl_ctx *ctx, int ticket, float *out)
{
    if (!ctx || !out || ticket < 0 || ticket >= pl_ctx::kReadbacks || !ctx->readback[ticket].busy)
        return pl_set_error(PL_ERR_ARG, "bad read-back ticket");
    pl_ctx::Readback &rb = ctx->readback[ticket];
    PL_CUDA(cudaSetDevice(ctx->device));
    PL_CUDA(cudaEventSynchronize(rb.done));
    memcpy(out, rb.pinned, sizeof(float2) * (size_t) rb.n);
    rb.busy = 0;
    return PL_OK;
}

static int check_norm_args(pl_ctx *ctx, const pl_norm_scene *sc, pl_pool *norm, pl_pool *elev, int n)
{
    if (!ctx || !sc || !norm || !elev) return pl_set_error(PL_ERR_ARG, "NULL argument");
    if (n < 0) return pl_set_error(PL_ERR_ARG, "n < 0");
    if (norm->kind != PL_POOL_NORM_UN8x2 && norm->kind != PL_POOL_NORM_UN8x4)
        return pl_set_error(PL_ERR_ARG, "norm is not a normal pool");
    if (elev->kind != PL_POOL_ELEV_F32x3) return pl_set_error(PL_ERR_ARG, "elev is not an elevation pool");
    if (norm->tile_w != sc->tile_w) return pl_set_error(PL_ERR_ARG, "scene tile_w != normal pool tile_w");
    if (elev->tile_w != sc->tile_w + 2 * sc->elev_border)
        return pl_set_error(PL_ERR_ARG, "elevation tile_w %d != normal tile_w %d + 2*border %d", elev->tile_w, sc->tile_w, sc->elev_border);
    if (sc->elev_border < 1) return pl_set_error(PL_ERR_ARG, "elev_border must be >= 1");
    return PL_OK;
}

extern "C" int pl_normal_batch_dev(pl_ctx *ctx, const pl_norm_scene *sc, pl_pool *norm, pl_pool *elev, int n,
                                   const pl_norm_req *dev_reqs)
{
    int rc = check_norm_args(ctx, sc, norm, elev, n);
    if (rc) return rc;
    if (n == 0) return PL_OK;
    if (!dev_reqs) return pl_set_error(PL_ERR_ARG, "reqs is NULL");
    return pl_launch_normal(ctx, sc, norm, elev, n, dev_reqs);
}

static int check_norm_reqs(const pl_pool *norm, const pl_pool *elev, int n, const pl_norm_req *reqs)
{
    for (int i = 0; i < n; ++i) {
        const pl_norm_req &q = reqs[i];
        if (q.out_slot < 0 || q.out_slot >= norm->capacity || q.elev_slot < 0 || q.elev_slot >= elev->capacity ||
            q.parent_slot >= norm->capacity || (q.parent_slot >= 0 && q.parent_slot == q.out_slot))
            return pl_set_error(PL_ERR_ARG, "request %d: slot out of range", i);
    }
    return PL_OK;
}

extern "C" int pl_normal_batch(pl_ctx *ctx, const pl_norm_scene *sc, pl_pool *norm, pl_pool *elev, int n,
                               const pl_norm_req *reqs)
{
    int rc = check_norm_args(ctx, sc, norm, elev, n);
    if (rc) return rc;
    if (n == 0) return PL_OK;
    if (!reqs) return pl_set_error(PL_ERR_ARG, "reqs is NULL");
    if ((rc = check_norm_reqs(norm, elev, n, reqs)) != PL_OK) return rc;
    void *dev = nullptr;
    rc = pl_stage_requests(ctx, reqs, sizeof(pl_norm_req) * (size_t) n, &dev);
    if (rc) return rc;
    return pl_launch_normal(ctx, sc, norm, elev, n, (const pl_norm_req *) dev);
}

/* ------------------------------------------------------------ tile pairs */


int pl_check_pair_pools(pl_ctx *ctx, const pl_elev_scene *esc, const pl_norm_scene *nsc, pl_pool *elev, pl_pool *norm, pl_pool *resid, int n)
{
    int rc = check_elev_args(ctx, esc, elev, resid, n);
    if (rc) return rc;
    return check_norm_args(ctx, nsc, norm, elev, n);
}

extern "C" int pl_pair_batch_dev(pl_ctx *ctx, const pl_elev_scene *esc, const pl_norm_scene *nsc, pl_pool *elev,
                                 pl_pool *norm, pl_pool *resid, int n, const pl_elev_req *dev_ereqs,
                                 const pl_norm_req *dev_nreqs)
{
    int rc = check_elev_args(ctx, esc, elev, resid, n);
    if (rc) return rc;
    if ((rc = check_norm_args(ctx, nsc, norm, elev, n)) != PL_OK) return rc;
    if (n == 0) return PL_OK;
    if (!dev_ereqs || !dev_nreqs) return pl_set_error(PL_ERR_ARG, "reqs is NULL");
    if (pl_pair_supported(ctx, esc, nsc, elev, norm))
        return pl_launch_pair(ctx, esc, nsc, elev, norm, resid, n, dev_ereqs, dev_nreqs);
    /* other geometries / RGBA8 normals: the two passes, one after the other */
    if ((rc = pl_launch_elevation(ctx, esc, elev, resid, n, dev_ereqs)) != PL_OK) return rc;
    return pl_launch_normal(ctx, nsc, norm, elev, n, dev_nreqs);
}

extern "C" int pl_pair_batch(pl_ctx *ctx, const pl_elev_scene *esc, const pl_norm_scene *nsc, pl_pool *elev,
                             pl_pool *norm, pl_pool *resid, int n, const pl_elev_req *ereqs, const pl_norm_req *nreqs)
{
    int rc = check_elev_args(ctx, esc, elev, resid, n);
    if (rc) return rc;
    if ((rc = check_norm_args(ctx, nsc, norm, elev, n)) != PL_OK) return rc;
    if (n == 0) return PL_OK;
    if (!ereqs || !nreqs) return pl_set_error(PL_ERR_ARG, "reqs is NULL");
    /* validate the requests and copy them into the pinned staging slot in one pass, on all host
     * threads for large batches (a planet level is 65 536 requests = 20 MB) */
    StageTicket tk;
    void *pe = nullptr, *pn = nullptr;
    rc = stage_acquire(ctx, sizeof(pl_elev_req) * (size_t) n, sizeof(pl_norm_req) * (size_t) n, &tk, &pe, &pn);
    if (rc) return rc;
    std::atomic<int> bad(n);
    std::atomic<int> not_reg(0);      /* some tile does not qualify for the register form (plnorm::normal_reg_ok, restated) */
    const bool sphere = nsc->sphere != 0;
    parallel_chunks(n, 4096, [&](int lo, int hi) {
        for (int i = lo; i < hi; ++i) {
            const pl_elev_req &e = ereqs[i];
            const pl_norm_req &q = nreqs[i];
            const bool reg = sphere ? q.smooth == 1.0f
                                    : (q.w2t[0] == 1.0f && q.w2t[1] == 0.0f && q.w2t[2] == 0.0f && q.w2t[3] == 0.0f && q.w2t[4] == 1.0f && q.w2t[5] == 0.0f);
            if (!reg) not_reg.store(1, std::memory_order_relaxed);
            int kind = 0;
            if (elev_req_bad(e, elev, resid))
                kind = 1;
            else if (q.out_slot < 0 || q.out_slot >= norm->capacity || q.elev_slot < 0 || q.elev_slot >= elev->capacity ||
                     q.parent_slot >= norm->capacity || (q.parent_slot >= 0 && q.parent_slot == q.out_slot))
                kind = 2;
            else if (q.elev_slot != e.out_slot)
                kind = 3;
            if (kind) {
                int cur = bad.load();
                while (i < cur && !bad.compare_exchange_weak(cur, i)) {}
                return;
            }
        }
        memcpy(static_cast<pl_elev_req *>(pe) + lo, ereqs + lo, sizeof(pl_elev_req) * (size_t) (hi - lo));
        memcpy(static_cast<pl_norm_req *>(pn) + lo, nreqs + lo, sizeof(pl_norm_req) * (size_t) (hi - lo));
    });
    if (bad.load() < n) {
        const int i = bad.load();
        /* re-derive the kind for the first bad request (threads may have raced on bad_kind) */
        const pl_elev_req &e = ereqs[i];
        if (elev_req_bad(e, elev, resid)) return pl_set_error(PL_ERR_ARG, "request %d: slot, noise index or window out of range", i);
        if (nreqs[i].elev_slot != e.out_slot && nreqs[i].elev_slot >= 0 && nreqs[i].elev_slot < elev->capacity &&
            nreqs[i].out_slot >= 0 && nreqs[i].out_slot < norm->capacity)
            return pl_set_error(PL_ERR_ARG, "request %d: the normal request does not read the elevation tile of the same index", i);
        return pl_set_error(PL_ERR_ARG, "request %d: slot out of range", i);
    }
    void *de = nullptr, *dn = nullptr;
    if ((rc = stage_commit(ctx, tk, &de, &dn)) != PL_OK) return rc;
    if (pl_pair_supported(ctx, esc, nsc, elev, norm))
        return pl_launch_pair(ctx, esc, nsc, elev, norm, resid, n, (const pl_elev_req *) de, (const pl_norm_req *) dn, not_reg.load() == 0);
    return pl_pair_batch_dev(ctx, esc, nsc, elev, norm, resid, n, (const pl_elev_req *) de, (const pl_norm_req *) dn);
}
