/*
 * pl_internal.h -- shared between the translation units of libproland_b200.so.
 * Not part of the C ABI (include/proland_b200.h is).
 */
#ifndef PL_INTERNAL_H
#define PL_INTERNAL_H

#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdarg.h>
#include <deque>
#include <vector>
#include "proland_b200.h"

/* ---- HBM layout of the pools (DESIGN.md "Data layout") ----------------------
 * ELEV_F32x3 : per slot 3 planes (zf, zc, zm); a plane is tile_w rows of `pitch`
 *              floats, pitch = tile_w rounded up to 4 (16-byte rows: TMA tensor
 *              maps need 16-byte strides, 101*4 is not).  Pad columns hold 0.
 * NORM_UN8xC : dense rows of tile_w*C bytes; slot stride rounded up so the
 *              last 16-byte bulk store of a tile stays inside its slot.
 * RESID_F32  : tile_w rows of pitch floats (pitch = tile_w rounded up to 4).
 * RESID_I16  : tile_w rows of pitch int16  (pitch = tile_w rounded up to 8).
 * ORTHO_UN8x4: dense rows of tile_w*4 bytes (16-byte rows: (tile_w-4) % 8 == 0), slot stride
 *              rounded up to 128.
 */
struct pl_pool {
    pl_ctx *ctx;
    int kind;
    int tile_w;
    int capacity;
    int pitch;          /* elements per row (floats / int16) or bytes per row (normals) */
    size_t plane_elems; /* ELEV: tile_w * pitch */
    size_t slot_bytes;
    size_t tile_bytes;  /* reference layout */
    uint8_t *base;      /* device */
    float2 *stats;      /* ELEV only: per-slot (zmin, zmax) */
    int *ready;         /* ELEV only: per-slot epoch of the pl_produce_levels call that finished the tile's elevation planes */
    /* peer copies of this pool on the other GPUs of the box (pl_pool_attach_peers): device pointers into
     * their memory, mapped through CUDA IPC; kernels that push finished tiles store to them over NVLink */
    enum { kMaxPeers = 7 };
    int npeers;
    int push;              /* NORM pools: the fused kernel also stores every tile into the peers (1: unicast, 2: multicast) */
    uint8_t *peer_base[kMaxPeers];
    /* NVLink multicast (pl_multicast.cu): the pool's memory is a VMM allocation bound to a multicast object shared by
     * the ranks of the box; a store to mc_base + offset lands at base + offset on EVERY GPU of the group */
    int vmm;                               /* base comes from cuMemCreate / cuMemMap (pl_pool_create_shared) */
    size_t vmm_size;
    unsigned long long vmm_handle;         /* CUmemGenericAllocationHandle of the pool's memory */
    unsigned long long mc_handle;          /* ... of the multicast object (0: none) */
    int mc_devices, mc_state;              /* 1: object known, 2: device added, 3: bound and mapped */
    uint8_t *mc_base;
    CUtensorMap tm_parent; /* ELEV only: 3-D map over the zf/zc/zm planes, box = parent window */
    int box_w, box_h;
};

struct pl_ctx {
    int device;
    int sm_count;
    cudaStream_t own_stream;
    cudaStream_t stream;
    uint64_t launches;
    uint64_t height_unsure;   /* pl_height_cube_from_latlon: base samples redone with the host's libm (pl_heights.cu) */
    /* noise (createDemNoise), 4 rotations x 6 layers of fp16: one table per tile width, as the reference's
     * demNoiseFactory caches one texture per width (ElevationProducer.cpp:135) -- producers of different tile sizes
     * share a context */
    struct NoiseTable { int w, pitch; __half *rot; };
    NoiseTable noise_tabs[8];
    int n_noise;
    const NoiseTable *noise_for(int w) const
    {
        for (int i = 0; i < n_noise; ++i)
            if (noise_tabs[i].w == w) return &noise_tabs[i];
        return nullptr;
    }
    /* ortho noise (createOrthoNoise), 4 rotations x 6 layers of RGBA8 */
    int ortho_noise_w;
    uint32_t *ortho_noise_rot;
    /* device-side request generation (pl_requests.cu) */
    int *perlin_perm;
    float *perlin_g2;
    pl_elev_req *gen_ereq;
    pl_norm_req *gen_nreq;
    int gen_cap;
    int force_generic;   /* tests: run the runtime-geometry kernels even for the shipped geometry */
    int no_fuse;         /* tests / profiling: pl_produce_range launches the two passes separately */
    int no_slim;         /* tests / profiling: the fused kernel never uses its slim (4 CTAs per SM) layout */
    int slim_sphere;     /* tests / profiling: ... and uses it on spheres too (pl_debug_no_slim(ctx, -1)) */
    int levels_epoch;    /* pl_produce_levels: the value a finished tile's ready flag takes in the current call */
    int inflate_path;    /* tests / profiling: 0 = by batch size, 1 = warp-per-stream decoder, 2 = tokenizer + resolver */
    /* per-launch CUDA-event timing (pl_timing_*): events on the launching stream */
    int timing;
    struct TimedLaunch { cudaEvent_t a, b; int kernel; int tiles; };
    std::vector<TimedLaunch> *timed;
    std::vector<cudaEvent_t> *event_pool;
    /* residual decode scratch (dense inflated streams + status) */
    void *resid_scratch;
    size_t resid_scratch_bytes;
    /* request staging */
    /* Request staging: ONE pinned ring + ONE device ring of the same size, sub-allocated in FIFO order.
     * An entry's pinned bytes are reused once its upload has finished (event, host wait), its device
     * bytes once the kernel that consumed it has finished (event, the copy stream waits).  The host can
     * run a whole root-to-leaf chain of batches ahead of the GPU. */
    struct StageEntry { size_t off, bytes; cudaEvent_t copied, consumed; int consumed_rec; cudaStream_t stream; };
    void *stage_dev, *stage_pinned;
    size_t stage_size, stage_w;
    size_t stage_min;            /* smallest ring to allocate (0: 32 MB); tests shrink it to force wrap-around */
    std::deque<StageEntry> *stage_fifo;
    std::vector<cudaEvent_t> *stage_events;   /* recycled (timing disabled: not interchangeable with event_pool) */
    cudaStream_t copy_stream;    /* request uploads run here, beside the kernels of earlier batches */
    /* asynchronous statistics read-backs (pl_elev_stats_readback_*) */
    struct Readback { void *pinned; size_t cap; cudaEvent_t done; int n; int busy; };
    enum { kReadbacks = 4 };
    Readback readback[kReadbacks];
    void *zgather;               /* device staging of pl_elev_zreadback_begin */
};

int pl_set_error(int code, const char *fmt, ...);
/* pl_multicast.cu: a pool's memory on the VMM allocator (pl_pool_create_shared) */
int pl_vmm_alloc(pl_pool *p, size_t bytes);
void pl_vmm_free(pl_pool *p);
#define PL_CUDA(call)                                                                       \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess)                                                              \
            return pl_set_error(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver \
                                    ? PL_ERR_NO_DEVICE : PL_ERR_CUDA,                       \
                                "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)

/* kernel ids of pl_timing_collect */
enum { PL_K_ELEVATION = 0, PL_K_NORMAL = 1, PL_K_GENREQ = 2, PL_K_RESIDUAL = 3, PL_K_PAIR = 4, PL_K_ORTHO = 5, PL_K_COUNT = PL_TIMING_KERNELS };
/* bracket a launch with events when timing is on: call begin before, end after */
void pl_timing_begin(pl_ctx *ctx, int kernel, int tiles);
void pl_timing_end(pl_ctx *ctx);

/* stage host requests on the device (returns device pointers; b may be NULL).  Both arrays travel in one
 * copy; nothing here waits for the GPU unless all staging slots are still in flight */
int pl_stage_requests2(pl_ctx *ctx, const void *a, size_t abytes, const void *b, size_t bbytes, void **adev, void **bdev);
static inline int pl_stage_requests(pl_ctx *ctx, const void *host, size_t bytes, void **dev)
{
    return pl_stage_requests2(ctx, host, bytes, nullptr, 0, dev, nullptr);
}

/* kernel launchers (defined next to their kernels) */
int pl_launch_elevation(pl_ctx *ctx, const pl_elev_scene *sc, pl_pool *elev, pl_pool *resid,
                        int n, const pl_elev_req *dev_reqs);
int pl_launch_normal(pl_ctx *ctx, const pl_norm_scene *sc, pl_pool *norm, pl_pool *elev,
                     int n, const pl_norm_req *dev_reqs);

int pl_check_pair_pools(pl_ctx *ctx, const pl_elev_scene *esc, const pl_norm_scene *nsc, pl_pool *elev, pl_pool *norm, pl_pool *resid, int n);
/* fused elevation + normal pass (pl_pair.cu): one CTA per tile pair */
bool pl_pair_supported(const pl_ctx *ctx, const pl_elev_scene *esc, const pl_norm_scene *nsc, const pl_pool *elev,
                       const pl_pool *norm);
int pl_launch_pair_levels(pl_ctx *ctx, const pl_elev_scene *esc, const pl_norm_scene *nsc, pl_pool *elev, pl_pool *norm, int n,
                           const pl_elev_req *dev_ereqs, const pl_norm_req *dev_nreqs, int epoch, bool all_reg = false);
/* all_reg: the caller knows that every tile of the launch qualifies for the register form of the FAST normal pass
 * (norm_level_all_reg / plnorm::normal_reg_ok): the fused kernel may use its slim layout (4 CTAs per SM) */
int pl_launch_pair(pl_ctx *ctx, const pl_elev_scene *esc, const pl_norm_scene *nsc, pl_pool *elev, pl_pool *norm,
                   pl_pool *resid, int n, const pl_elev_req *dev_ereqs, const pl_norm_req *dev_nreqs, bool all_reg = false);

int pl_launch_ortho(pl_ctx *ctx, const pl_ortho_scene *sc, pl_pool *ortho, pl_pool *resid, int n,
                    const pl_ortho_req *dev_reqs);
void pl_host_ortho_noise(int W, uint8_t *out6);   /* createOrthoNoise, 6*W*W*4 bytes */

/* host maths shared with the request builders (pl_hostmath.cpp) */
void pl_host_dem_noise(int W, float *out6);      /* fp32, before the R16F rounding */

static inline int pl_round_up(int v, int m) { return (v + m - 1) / m * m; }

/* Column of texel x inside a row of the pre-rotated noise planes.  Each group of four texels is
 * stored as (x0, x0+2, x0+1, x0+3): the elevation kernel works on the two horizontally adjacent
 * 2x2 quads of a group with packed fp32x2 maths, lane 0 = first quad, lane 1 = second, so one
 * 8-byte load yields the half2 (even texel of both quads), then the half2 (odd texel of both). */
#if defined(__CUDACC__)
__host__ __device__
#endif
static inline int pl_noise_col(int x) { return (x & ~3) | ((x & 1) << 1) | ((x >> 1) & 1); }

#endif
