/*
 * proland_b200.h -- C ABI of the B200-native terrain tile-production path
 * (ElevationProducer -> NormalProducer, fed by ResidualProducer; and its colour
 * twin OrthoProducer, fed by OrthoCPUProducer).
 *
 * Plain C, plain pointers and sizes.  This is the boundary a Proland build
 * binds instead of its GLSL draw calls; INTEGRATION.md shows the C++ side.
 * Reference paths below are relative to the reference checkout.
 *
 * Every function returns a pl_status (0 = ok) unless stated otherwise.  The
 * library never falls back to the CPU: without a CUDA device every entry point
 * that needs one returns PL_ERR_NO_DEVICE.
 */
#ifndef PROLAND_B200_H
#define PROLAND_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PL_ABI_VERSION 2

/* Error convention: the reference asserts / logs / returns NULL
 * (TileSampler.cpp:441-444, ResidualProducer.cpp:86-91); here each
 * condition has a code and pl_last_error() gives the text that would be logged. */
typedef enum pl_status {
    PL_OK = 0,
    PL_ERR_ARG = 1,        /* bad argument (an assert in the reference)          */
    PL_ERR_POOL_FULL = 2,  /* "Insufficient tile cache size"                     */
    PL_ERR_CUDA = 3,       /* a CUDA runtime / driver call failed                */
    PL_ERR_CORRUPT = 4,    /* residual container / blob does not parse           */
    PL_ERR_NO_DEVICE = 5,  /* no CUDA device: there is no CPU fallback           */
    PL_ERR_IO = 6          /* residual file cannot be opened (maxLevel = -1)     */
} pl_status;

const char *pl_last_error(void);
int pl_abi_version(void);

/* ------------------------------------------------------------------ context */

typedef struct pl_ctx pl_ctx;
typedef struct pl_pool pl_pool;

/* One context per GPU / per producer thread (replaces the GL context + FBO of
 * ElevationProducer.cpp:136-155).  Work is ordered on the context's stream. */
int pl_ctx_create(int device, pl_ctx **out);
void pl_ctx_destroy(pl_ctx *ctx);
/* Use an existing cudaStream_t (e.g. torch's current stream) instead of the
 * context's own; NULL restores the own stream. */
int pl_ctx_set_stream(pl_ctx *ctx, void *cuda_stream);
void *pl_ctx_stream(pl_ctx *ctx);
int pl_sync(pl_ctx *ctx);
/* kernels launched by this context since creation (bench.py's gpu_launches) */
uint64_t pl_ctx_launch_count(const pl_ctx *ctx);
int pl_device_sm_count(pl_ctx *ctx);

/* Per-launch timing with CUDA events recorded on the launching stream (the
 * reference's hook is Ork's monitorTask("CreateElevationTile")).
 * pl_timing_collect synchronises, sums the elapsed time per kernel
 * {0 elevation, 1 normal, 2 request generation, 3 residual decode, 4 fused
 * elevation+normal, 5 ortho} into the three PL_TIMING_KERNELS-entry arrays and
 * resets the record. */
#define PL_TIMING_KERNELS 6
int pl_timing_enable(pl_ctx *ctx, int on);
int pl_timing_collect(pl_ctx *ctx, double *ms, uint64_t *launches, uint64_t *tiles);

/* ------------------------------------------------- tile pools (TileStorage) */

/* Replaces GPUTileStorage (producer/GPUTileStorage.cpp:129-199) and, for
 * residuals, the CPUTileStorage<float> slots ElevationProducer reads
 * (ElevationProducer.cpp:326-337): one device slab, a slot is an index.
 * Slot bookkeeping (free list, LRU) stays on the host (TileStorage/TileCache). */
typedef enum pl_pool_kind {
    PL_POOL_ELEV_F32x3 = 0, /* RGB32F elevation (zf,zc,zm); planar, padded pitch  */
    PL_POOL_NORM_UN8x2 = 1, /* RG8 normals                                        */
    PL_POOL_NORM_UN8x4 = 2, /* RGBA8 normals (fine + coarse)                      */
    PL_POOL_RESID_F32 = 3,  /* float residual tiles (CPUTileStorage<float>)       */
    PL_POOL_RESID_I16 = 4,  /* raw int16 residual tiles as stored in the file     */
    PL_POOL_ORTHO_UN8x4 = 5 /* RGBA8 ortho tiles and their byte residual tiles
                               (OrthoProducer's GPUTileStorage / the CPUTileStorage<unsigned char>
                               slots it uploads, ortho/OrthoProducer.cpp:296-318); a storage with
                               fewer channels keeps the first ones                  */
} pl_pool_kind;

int pl_pool_create(pl_ctx *ctx, int kind, int tile_w, int capacity, pl_pool **out);
void pl_pool_destroy(pl_pool *pool);
int pl_pool_capacity(const pl_pool *pool);
int pl_pool_tile_w(const pl_pool *pool);
/* bytes of one tile in the REFERENCE's layout (what download/upload move):
 * tile_w*tile_w*{12, 2, 4, 4, 2, 4} */
size_t pl_pool_tile_bytes(const pl_pool *pool);
/* bytes of one slot in HBM (padded) and the device base pointer */
size_t pl_pool_slot_bytes(const pl_pool *pool);
void *pl_pool_device_ptr(pl_pool *pool);
/* copy one slot to / from host memory in the reference's layout: interleaved
 * (zf,zc,zm) floats, row-major, row 0 first (GPUSlot::copyPixels order);
 * interleaved bytes for normals; dense floats / int16 for residuals. */
int pl_pool_download(pl_pool *pool, int slot, void *host, size_t bytes);
int pl_pool_upload(pl_pool *pool, int slot, const void *host, size_t bytes);
/* n consecutive slots, same layout per tile, one device-to-host copy (bytes = n * pl_pool_tile_bytes) */
int pl_pool_download_range(pl_pool *pool, int slot0, int n, void *host, size_t bytes);

/* Peers: the copies of a pool on the other GPUs of one box (one process per GPU).  pl_pool_export gives the
 * PL_IPC_HANDLE_BYTES-byte CUDA IPC handle of the pool's memory; after the processes have exchanged them (any
 * transport: torch.distributed.all_gather in tools/gather_tiles.py) pl_pool_attach_peers(handles of all n ranks in
 * rank order, own rank) maps the others' pools (peer access over NVLink).  With pl_pool_push_to_peers(norm, 1) the
 * fused elevation + normal kernel stores every finished RG8 normal tile into the same slot of every peer's pool
 * as well: the "gather finished tiles" step of a multi-GPU sweep rides on the producing kernel, tile
 * by tile, instead of a collective after it.  Slot numbering must be the same on all ranks; a consumer reads the
 * tiles after the producer's stream has been synchronised and the ranks have met (a barrier).  All pools of the
 * group must have the same kind, tile_w and capacity. */
#define PL_IPC_HANDLE_BYTES 64
int pl_pool_export(pl_pool *pool, void *handle64);
int pl_pool_attach_peers(pl_pool *pool, int n, const void *handles, int self);
int pl_pool_push_to_peers(pl_pool *pool, int on);

/* The same push through ONE store: an NVLink multicast object (cuMulticastCreate) every rank binds its pool's memory to;
 * a store to the object's mapping is replicated by the NVSwitch into the same offset of every bound pool, so one copy of a
 * tile leaves the producing GPU instead of one per peer.  The pools come from pl_pool_create_shared (VMM allocator, same
 * kind / tile_w / capacity on every rank; otherwise identical to pl_pool_create).  Order, one process per GPU:
 *   rank 0: pl_pool_mc_create(pool, n, &fd) -- fd is a POSIX file descriptor naming the object; the CALLER hands it to the
 *           other ranks (SCM_RIGHTS over a Unix socket; the library does no inter-process transport) -- the others:
 *           pl_pool_mc_import(pool, their copy of fd, n); every rank: pl_pool_mc_add_device; BARRIER; every rank:
 *           pl_pool_mc_bind; BARRIER; pl_pool_push_to_peers(pool, 2) selects the multicast push (1: unicast, 0: off).
 * The file descriptors may be closed after the import.  PL_ERR_ARG when the device has no multicast support. */
int pl_pool_create_shared(pl_ctx *ctx, int kind, int tile_w, int capacity, pl_pool **out);
int pl_pool_mc_create(pl_pool *pool, int n_devices, int *fd_out);
int pl_pool_mc_import(pl_pool *pool, int fd, int n_devices);
int pl_pool_mc_add_device(pl_pool *pool);
int pl_pool_mc_bind(pl_pool *pool);

/* ------------------------------------------------------------------- noise */

/* createDemNoise (ElevationProducer.cpp:50-133): builds the six W x W layers
 * with the reference's LCG, rounds to fp16 (the R16F upload) and stores the 4
 * rotations of each layer on the device.  host_out (optional, 6*W*W floats)
 * receives the fp16-rounded layers. */
int pl_noise_init(pl_ctx *ctx, int tile_w, float *host_out);
/* ElevationProducer.cpp:345-373: noise layer / rotation of a tile (host). */
void pl_noise_select(int level, int tx, int ty, int face, int *noiseR, int *noiseL);
/* noise.cpp:117-165 (2D) -- exported so the host mirror and tests share it */
float pl_cnoise2(float x, float y);

/* --------------------------------------------------------------- elevation */

enum { PL_NOISE_PLAIN = 0,   /* upsampleShader variants A, C: zf += |rs|*n        */
       PL_NOISE_SLOPE = 1 }; /* variants B, D: slope / curvature modulated        */

/* per-producer constants = the tile-independent uniforms + shader variant */
typedef struct pl_elev_scene {
    int32_t tile_w;      /* tileWSDF.x, e.g. 101                                  */
    int32_t grid;        /* tileWSDF.z = (tile_w-5)/gridSize                      */
    int32_t flip;        /* tileWSDF.w (and the variant honours it)               */
    int32_t noise_mode;  /* PL_NOISE_*                                            */
    int32_t no_clamp;    /* #define NO_CLAMP (upsampleShader-noClamp.xml)         */
    int32_t want_stats;  /* also write per-slot (zmin,zmax) of zm, TileSamplerZ   */
    float resid_scale;   /* int16 -> metres factor when the residual pool is I16
                            (ResidualProducer.cpp:333: the file's scale * zscale) */
    int32_t pad_;
} pl_elev_scene;

/* per-tile uniforms, ElevationProducer.cpp:305-376 */
typedef struct pl_elev_req {
    int32_t out_slot;     /* GPUSlot::l of the tile being produced                */
    int32_t parent_slot;  /* coarseLevelOSL.w, -1 at level 0                      */
    int32_t resid_slot;   /* slot in the residual pool, -1: residualOSH.w = 0     */
    int32_t dx, dy;       /* coarseLevelOSL.xy in texels: (t%2)*(tile_w-5)/2      */
    int32_t rx, ry;       /* residual window origin, ElevationProducer.cpp:324    */
    int32_t noise_r;      /* noiseUVLH.x                                          */
    int32_t noise_l;      /* noiseUVLH.z                                          */
    float rs;             /* noiseUVLH.w                                          */
    float pixel_size;     /* tileWSDF.y                                           */
    int32_t level, tx, ty;/* informational (kept for device-generated batches)    */
    int32_t pad_[2];
} pl_elev_req;            /* 64 bytes */

/* Fill one request exactly as ElevationProducer::doCreateTile does (host). */
void pl_elev_make_req(int tile_w, float root_quad_size, const float *noise_amp, int n_amp,
                      int face, int level, int tx, int ty, int resid_tile_w, int has_resid,
                      pl_elev_req *req);

/* Requests handed over as HOST arrays are validated (slots, noise indices, dx / dy in {0, tileSize / 2}, the residual
 * window inside its tile with rx a multiple of 4); the *_dev variants trust the caller.  A batch must not contain both a
 * tile and its parent (or a tile whose parent_slot is another request's out_slot): tiles of a launch are produced
 * concurrently -- produce level by level, as TileProducer's task graph orders a tile behind its parent. */
/* The batched upsampleShader: n tiles, one CTA per tile.  reqs is HOST memory
 * (copied to the device inside the call).  resid may be NULL. */
int pl_elevation_batch(pl_ctx *ctx, const pl_elev_scene *scene, pl_pool *elev,
                       pl_pool *resid, int n, const pl_elev_req *reqs);
/* same, requests already on the device */
int pl_elevation_batch_dev(pl_ctx *ctx, const pl_elev_scene *scene, pl_pool *elev,
                           pl_pool *resid, int n, const pl_elev_req *dev_reqs);
/* per-slot (zmin,zmax) written by want_stats batches; out = 2*n floats */
int pl_elev_stats_download(pl_ctx *ctx, pl_pool *elev, int n, const int32_t *slots, float *out);
/* same for the contiguous slots [slot0, slot0+n) -- the readback TileSamplerZ does
 * (core/sources/proland/terrain/TileSamplerZ.cpp:253-351), 8 bytes per tile */
int pl_elev_stats_range(pl_ctx *ctx, pl_pool *elev, int slot0, int n, float *out);
/* The same read-back, asynchronous: _begin enqueues the copy behind the work already on the stream and
 * returns a ticket, _end waits for that copy only and delivers n (zmin, zmax) pairs.  This is how the
 * reference consumes the values: TileSamplerZ reads them back through ReadbackManager a few frames late
 * (TileSamplerZ.cpp:253-351).  At most 4 read-backs can be in flight. */
int pl_elev_stats_readback_begin(pl_ctx *ctx, pl_pool *elev, int slot0, int n, int *ticket);
int pl_elev_stats_readback_end(pl_ctx *ctx, int ticket, float *out);
/* TileSamplerZ's frame read-back (core/sources/proland/terrain/TileSamplerZ.cpp:253-351): the z range of up to 64 tiles in
 * arbitrary slots and, when cam_slot >= 0, the zm texel (cam_x, cam_y) of that slot -- the ground height under the
 * camera.  Collected with pl_elev_stats_readback_end: [(h, h) of the camera texel], then (zmin, zmax) per slot.
 * pl_elev_stats_readback_ready: 1 once the result has arrived (the collection will not wait), 0 before. */
int pl_elev_zreadback_begin(pl_ctx *ctx, pl_pool *elev, int n, const int32_t *slots, int cam_slot, int cam_x, int cam_y,
                            int *ticket);
int pl_elev_stats_readback_ready(pl_ctx *ctx, int ticket);

/* ----------------------------------------------------------------- normals */

enum { PL_FILTER_NEAREST = 0, PL_FILTER_LINEAR = 1 };

typedef struct pl_norm_scene {
    int32_t tile_w;        /* tileSDF.x, e.g. 97                                  */
    int32_t grid;          /* tileSDF.y                                           */
    int32_t elev_border;   /* elevationTiles->getBorder() = 2                     */
    int32_t elev_filter;   /* min/mag filter of the elevation storage             */
    int32_t parent_filter; /* min/mag filter of the normal storage                */
    int32_t sphere;        /* deform="sphere"                                     */
    int32_t arith;         /* PL_ARITH_EXACT (0, default) or PL_ARITH_FAST: see below */
    int32_t pad_;
} pl_norm_scene;
/* The arithmetic contract of the NORMAL pass (elevations are always bit-exact: children are built from them).
 *   PL_ARITH_EXACT  every operation of normalShader.glsl in the canonical fp32 order with IEEE division and square
 *                   root: tiles are bit-identical to the oracle.
 *   PL_ARITH_FAST   the tolerance contract: a normal byte differs from the canonical evaluation by at most ONE
 *                   unorm8 step (0.45 degrees at worst; BASELINE's "within a stated angular tolerance"), and fewer
 *                   than 1e-3 of the bytes of a tile set differ.  MUFU.RCP / MUFU.RSQ seeds, one reciprocal per grid
 *                   point, bilinear forms evaluated by row (pl_normal_tile.cuh).  Levels below R/64 quad size
 *                   (smoothstep < 1) keep the exact position code.  Served by the specialised kernels (97-texel
 *                   RG8 tiles); other geometries ignore it and stay exact. */
enum { PL_ARITH_EXACT = 0, PL_ARITH_FAST = 1 };

/* per-tile uniforms, NormalProducer.cpp:196-283 (fp64 on the host -> fp32) */
typedef struct pl_norm_req {
    int32_t out_slot;
    int32_t elev_slot;
    int32_t parent_slot;   /* parent normal tile, -1: normalOSL = -1              */
    int32_t ptx, pty;      /* tx%2, ty%2                                          */
    int32_t level;
    float deform[4];       /* x0, y0, quad size, R (0 = flat)                     */
    float corners[12];     /* patchCorners rows x,y,z (row w is 1,1,1,1)          */
    float verticals[12];   /* patchVerticals rows x,y,z (row w is 0)              */
    float norms[4];        /* patchCornerNorms                                    */
    float w2t[9];          /* worldToTangentFrame, row-major                      */
    float p2t[9];          /* parentToTangentFrame, row-major                     */
    float smooth;          /* smoothstep(R/32, R/64, deform.z)                    */
    int32_t pad_[3];
} pl_norm_req;             /* 240 bytes */

void pl_norm_make_req(const pl_norm_scene *scene, double root_quad_size, int components,
                      int level, int tx, int ty, pl_norm_req *req);

int pl_normal_batch(pl_ctx *ctx, const pl_norm_scene *scene, pl_pool *norm, pl_pool *elev,
                    int n, const pl_norm_req *reqs);
int pl_normal_batch_dev(pl_ctx *ctx, const pl_norm_scene *scene, pl_pool *norm, pl_pool *elev,
                        int n, const pl_norm_req *dev_reqs);

/* ------------------------------------------------ device-side batch driver */

/* Everything a producer pair (elevationProducer + normalProducer resources)
 * holds that does not depend on the tile. */
/* ------------------------------------------------- tile pairs (both passes) */

/* An elevation tile and the normal tile derived from it, produced by ONE kernel (pl_pair.cu): what the
 * reference runs as CreateElevationTile followed by CreateNormalTile on the same (level, tx, ty)
 * (TileProducer.cpp:199-217 -> ElevationProducer.cpp:280-405, NormalProducer.cpp:164-289).  Request i of
 * both arrays describes tile i: nreqs[i].elev_slot == ereqs[i].out_slot.  Results are bit-identical to
 * pl_elevation_batch followed by pl_normal_batch; geometries the fused kernel does not cover (tile
 * sizes other than 101/97) run as those two passes.  RGBA8 normal pools (NORM_UN8x4: fine + the parent's coarse normal,
 * normalShader.glsl:100-114) are served by the fused kernel in exact arithmetic: nreqs[i].parent_slot names the parent's
 * NORMAL tile, which -- like the parent elevation tile -- must not be produced by the same launch. */
int pl_pair_batch(pl_ctx *ctx, const pl_elev_scene *escene, const pl_norm_scene *nscene, pl_pool *elev,
                  pl_pool *norm, pl_pool *resid, int n, const pl_elev_req *ereqs, const pl_norm_req *nreqs);
int pl_pair_batch_dev(pl_ctx *ctx, const pl_elev_scene *escene, const pl_norm_scene *nscene, pl_pool *elev,
                      pl_pool *norm, pl_pool *resid, int n, const pl_elev_req *dev_ereqs,
                      const pl_norm_req *dev_nreqs);

/* A tile pair by its identity: what a producer KNOWS about the tile it has to create -- its quadtree coordinates and
 * the slots TileCache handed out -- in 32 bytes.  The uniforms ElevationProducer::doCreateTile / NormalProducer::
 * doCreateTile derive from them (ElevationProducer.cpp:305-376, NormalProducer.cpp:196-283: noise layer and rotation
 * through cnoise, parent window, residual window, pixel size; the fp64 patch geometry of the sphere) are expanded ON
 * THE DEVICE by the code pl_produce_range uses, so 32 instead of 304 bytes per tile cross PCIe and the host does no
 * per-tile arithmetic at all. */
typedef struct pl_tile_id {
    int32_t level, tx, ty;
    int32_t elev_slot;     /* output slot in the elevation pool                               */
    int32_t parent_slot;   /* the parent's slot in the elevation pool; ignored at level 0     */
    int32_t resid_slot;    /* slot in the residual pool, -1: the tile has no residual         */
    int32_t norm_slot;     /* output slot in the normal pool                                   */
    int32_t pad_;
} pl_tile_id;              /* 32 bytes */

typedef struct pl_sweep_scene {
    pl_elev_scene elev;
    pl_norm_scene norm;
    float root_quad_size;   /* TileProducer::getRootQuadSize()                      */
    int32_t face;           /* ElevationProducer.cpp:503-509                        */
    int32_t n_amp;          /* noise="..." list (<= 32 levels)                      */
    int32_t pad_;
    float noise_amp[32];
} pl_sweep_scene;

/* Produce n tiles of one level, consecutive in Morton order from morton0 (the
 * order TileSampler::getTiles visits children: x in the even bits), into slots
 * out_slot0 .. out_slot0+n-1 of BOTH pools (norm may be NULL: elevation only).
 * The parent of tile m is slot parent_slot0 + ((m >> 2) - parent_morton0) of
 * the elevation pool.  The per-tile uniforms are generated on the device; no
 * per-tile data crosses the PCIe bus.  Fractal scenes only (no residuals). */
int pl_produce_range(pl_ctx *ctx, const pl_sweep_scene *scene, pl_pool *elev, pl_pool *norm,
                     int level, uint64_t morton0, int n, int out_slot0, int parent_slot0,
                     uint64_t parent_morton0);
/* n tile pairs given by identity (ids: HOST memory, copied inside the call): the uniforms are generated on the device,
 * then the pair runs exactly as pl_pair_batch would on host-built requests (bit-identical results;
 * tests/test_gpu_sweep.py).  resid may be NULL; a tile's residual window is (tx % mod, ty % mod) * tileSize of its
 * residual tile (ElevationProducer.cpp:322-335).  A batch must not contain a tile together with its parent. */
int pl_pair_batch_ids(pl_ctx *ctx, const pl_sweep_scene *scene, pl_pool *elev, pl_pool *norm, pl_pool *resid,
                      int n, const pl_tile_id *ids);
/* Several consecutive levels of a subtree in ONE launch: the chain root -> leaves of a small sweep is otherwise a
 * sequence of launches of 1, 4, 16, .. tiles, each waiting for the level before it to drain (config 4, d = 6: 30
 * launches for 5 469 pairs).  Here the tiles of all ranges are CTAs of one grid, ordered by level; a tile of range
 * k >= 1 waits, inside the kernel, for the ready flag its parent (a tile of range k - 1, a CTA with a lower index) sets
 * once its elevation planes are stored -- children start while the parent still computes its normals.
 * Contract: ranges[k].level == ranges[k-1].level + 1, every parent of ranges[k] is a tile of ranges[k-1]
 * (parent_slot0 / parent_morton0 of range k are out_slot0 / morton0 of range k - 1); the parents of ranges[0] are
 * finished tiles.  Each range is laid out like a pl_produce_range call; results are bit-identical to those calls. */
typedef struct pl_level_range {
    int32_t level, n;
    uint64_t morton0;
    int32_t out_slot0, parent_slot0;
    uint64_t parent_morton0;
} pl_level_range;
int pl_produce_levels(pl_ctx *ctx, const pl_sweep_scene *scene, pl_pool *elev, pl_pool *norm, int nranges,
                      const pl_level_range *ranges);
/* The identities of a Morton range laid out like pl_produce_range's (host helper: what a caller walking the quadtree
 * in Morton order hands to pl_pair_batch_ids; the slot of tile m is out_slot0 + (m - morton0) in both pools) */
int pl_make_tile_ids_range(int level, uint64_t morton0, int n, int out_slot0, int parent_slot0,
                           uint64_t parent_morton0, pl_tile_id *ids);
/* The same requests built on the host with nthreads threads (<= 0: all): feeds
 * pl_elevation_batch / pl_normal_batch, and checks the device generator. */
int pl_make_requests_range(const pl_sweep_scene *scene, int level, uint64_t morton0, int n,
                           int out_slot0, int parent_slot0, uint64_t parent_morton0,
                           pl_elev_req *elev_reqs, pl_norm_req *norm_reqs, int nthreads);
/* test hooks: run the runtime-geometry kernels even for the shipped geometry;
 * evaluate the branch-free div / rcp / sqrt next to the IEEE operators on the
 * device: out = 6*n floats (div_rn, a/b, rcp_rn, 1/b, sqrt_rn(|a|), sqrtf(|a|)) */
int pl_debug_force_generic(pl_ctx *ctx, int on);
/* tests / profiling: make pl_produce_range and pl_pair_batch[_dev] launch the elevation and the normal
 * pass as two kernels instead of the fused one (same results, bit for bit) */
int pl_debug_no_fuse(pl_ctx *ctx, int on);
/* tests / profiling: the fused kernel never uses its slim layout (launches whose tiles all take the register form of the
 * FAST normal pass: 4 CTAs per SM; by default flat scenes only, where it pays); on = -1: use it on spheres too.  Results are
 * the same bit for bit */
int pl_debug_no_slim(pl_ctx *ctx, int on);
/* tests / profiling: which DEFLATE decoder pl_residual_decode_batch / pl_ortho_decode_batch run: 0 = chosen by the batch
 * size (default), 1 = the warp-per-stream kernel (small batches), 2 = the tokenizer + resolver pair (large batches) */
int pl_debug_inflate_path(pl_ctx *ctx, int path);
/* tests: the smallest request staging ring to allocate (default 32 MB); a small ring makes the FIFO
 * wrap around and grow within a few batches */
int pl_debug_stage_ring(pl_ctx *ctx, size_t min_bytes);
int pl_debug_fpexact(pl_ctx *ctx, int n, const float *a, const float *b, float *out);
/* copy the requests the last pl_produce_range generated back to the host (tests) */
int pl_debug_download_requests(pl_ctx *ctx, int n, pl_elev_req *elev_reqs, pl_norm_req *norm_reqs);

/* --------------------------------------------------------------- residuals */

/* ResidualProducer::readTile (ResidualProducer.cpp:268-340): n blobs, each a
 * little-endian TIFF with one DEFLATE strip of w*w little-endian int16.
 * blobs/offsets/sizes are HOST memory; tile j is written at the lower-left of
 * slot out_slots[j] of an I16 pool (raw) or an F32 pool (int16 * scale, added
 * to what add_slots[j] of the same pool holds when add_slots != NULL and
 * add_slots[j] >= 0).  widths[j] = w of tile j (<= pool tile_w). */
int pl_residual_decode_batch(pl_ctx *ctx, pl_pool *out, int n, const uint8_t *blobs,
                             const uint64_t *offsets, const uint32_t *sizes,
                             const int32_t *widths, const int32_t *out_slots,
                             const int32_t *add_slots, float scale);

/* A residual archive resident in device memory: the whole .dat file uploaded once (the reference maps it into the
 * address space once, ResidualProducer.cpp:70-129), tiles decoded from there.  pl_residual_decode_stored is
 * pl_residual_decode_batch without the per-batch packing and upload of the blobs: offsets / sizes locate each tile's
 * TIFF blob inside the archive, host_bytes is the caller's own copy (mapping) of the same file -- the few IFD
 * bytes of each blob are parsed on the host, the compressed strip is read on the device where it lies. */
typedef struct pl_blobs pl_blobs;
int pl_blobs_create(pl_ctx *ctx, const uint8_t *bytes, uint64_t size, pl_blobs **out);
void pl_blobs_destroy(pl_blobs *blobs);
int pl_residual_decode_stored(pl_ctx *ctx, pl_pool *out, const pl_blobs *store, const uint8_t *host_bytes, int n,
                              const uint64_t *offsets, const uint32_t *sizes, const int32_t *widths,
                              const int32_t *out_slots, const int32_t *add_slots, float scale);

/* A slot argument of the residual entry points may be PL_SLOT_SCRATCH: the hidden extra slot every
 * F32 residual pool carries (the `tmp` array of ResidualProducer::doCreateTile,
 * ResidualProducer.cpp:219). */
#define PL_SLOT_SCRATCH (-2)

/* ResidualProducer::upsample (ResidualProducer.cpp:342-384): dst = the (tile_size + 5)^2 tile of root
 * level `level` (tile_size = getTileSize(level)) upsampled from the quadrant (tx%2, ty%2) of the parent
 * tile in src, in the reference's CPU evaluation order.  Both slots are in the same F32 pool. */
int pl_residual_upsample(pl_ctx *ctx, pl_pool *pool, int src_slot, int dst_slot, int tile_size,
                         int tx, int ty);

/* ---- the step before the path: building residual files ----
 * One level of HeightMipmap::buildResiduals (preprocess/terrain/HeightMipmap.cpp:255-324) for n tiles:
 *   residual      = heights - upsample(parent approximation)        computeResidual   :449-497
 *   stored int16  = short(roundf(residual))                          encodeResidual    :499-512
 *   approximation = upsample(parent approximation) + stored int16    computeApproxTile :514-559
 * The upsample is ResidualProducer::upsample's (same taps, same CPU evaluation order), so the
 * approximation equals, bit for bit, what the run-time producer reconstructs from the file.
 * heights / approx: F32 residual pools (tiles of tile_size + 5 texels in the lower-left corner of the
 * slot, heights already divided by the file's scale); resid: an I16 residual pool.  parent_slot = -1 is the
 * level-0 rule (produceTile :567-578, getApproxTile :420-431): stored int16 = short(roundf(heights)), and the
 * approximation is the height tile itself, unrounded.
 * max_residual / max_err (n floats each, or NULL): max |residual before rounding|, max |heights - approximation|. */
typedef struct pl_resid_enc_req {
    int32_t tile_slot;      /* heights pool: the tile to encode                    */
    int32_t parent_slot;    /* approx pool: the parent's approximation; -1: level 0 */
    int32_t approx_slot;    /* approx pool: receives this tile's approximation     */
    int32_t resid_slot;     /* resid pool: receives the int16 residuals            */
    int32_t tile_size;      /* min(topLevelSize << level, tileSize)                */
    int32_t tx, ty;         /* the tile's coordinates (select the parent quadrant) */
    int32_t pad_;
} pl_resid_enc_req;         /* 32 bytes */
int pl_residual_encode_batch(pl_ctx *ctx, pl_pool *heights, pl_pool *approx, pl_pool *resid, int n,
                             const pl_resid_enc_req *reqs, float *max_residual, float *max_err);

/* ---- the height pyramid of the residual builder (the step before pl_residual_encode_batch) ------------------
 * HeightMipmap's base level, mipmap levels and cube-face stitching (preprocess/terrain/HeightMipmap.cpp:67-81, 149-254,
 * 327-372, 404-412; Preprocess.cpp:512-585 preprocessSphericalDem).  The six base-level grids ((base_size + 1)^2 int16
 * samples each, row-major; base_size = top_level_size << maxLevel) stay resident on the device; a height tile of any
 * level is gathered from them: level l = the base level decimated, samples near a cube corner collapse onto the corner,
 * samples past an edge come from the neighbouring face under setCube's rotation (nfaces = 6), or are clamped
 * (nfaces = 1: a flat DEM, AbstractTileCache.cpp:73-92).
 *   pl_height_cube_create       from host grids (faces[0..nfaces): hm1..hm6 of setCube)
 *   pl_height_cube_from_latlon  SphericalHeightFunction (Preprocess.cpp:406-445): the six grids from an equirectangular
 *                               source map (src_w x src_h floats, bilinear in lon / lat), computed on the device
 *   pl_height_cube_from_plane   PlaneHeightFunction (Preprocess.cpp:335-366): the one grid of a flat DEM (preprocessDem)
 *   pl_height_tiles             HeightMipmap::getTile for n (face, level, tx, ty): (ts + 5)^2 heights / scale into slot
 *                               out_slot of an F32 residual pool, ts = min(top_level_size << level, tile_size) */
typedef struct pl_height_cube pl_height_cube;
typedef struct pl_height_req {
    int32_t face, level, tx, ty;
    int32_t out_slot;
    int32_t pad_[3];
} pl_height_req;             /* 32 bytes */
int pl_height_cube_create(pl_ctx *ctx, int base_size, int nfaces, const int16_t *const *faces, pl_height_cube **out);
int pl_height_cube_from_latlon(pl_ctx *ctx, int base_size, const float *src, int src_w, int src_h, pl_height_cube **out);
int pl_height_cube_from_plane(pl_ctx *ctx, int base_size, const float *src, int src_w, int src_h, pl_height_cube **out);
/* how many base samples pl_height_cube_from_latlon has recomputed with the host's libm so far (the samples whose int16
 * truncation the device's atan2 / acos could not decide) */
uint64_t pl_debug_height_unsure(const pl_ctx *ctx);
int pl_height_cube_download(pl_ctx *ctx, const pl_height_cube *cube, int face, int16_t *out);
void pl_height_cube_destroy(pl_height_cube *cube);
int pl_height_tiles(pl_ctx *ctx, const pl_height_cube *cube, pl_pool *heights, int top_level_size, int tile_size,
                    float scale, int n, const pl_height_req *reqs);

/* HeightMipmap::generate + produceTile (HeightMipmap.cpp:99-130, 561-655): write a residual file in the
 * format ResidualProducer reads -- header, offset table per tile id, one little-endian TIFF/DEFLATE blob
 * per tile (levels below min_level first, then Lebesgue order per level), all-zero tiles sharing one
 * blob.  tiles + tile_offsets[id]: the dense (w x w, w = getTileSize(level) + 5) int16 residuals of tile
 * id (ResidualProducer::getTileId numbering).  zlib_level: -1 (zlib's default, what libtiff uses) .. 9.
 * Host code only; PL_ERR_IO when the file cannot be written. */
int pl_residual_write_file(const char *path, int min_level, int max_level, int tile_size, int root_level,
                           int root_tx, int root_ty, float scale, const int16_t *tiles,
                           const uint64_t *tile_offsets, int zlib_level);

/* ------------------------------------------------------------------- ortho
 * OrthoProducer: the colour twin of the elevation pass.  A tile is the 4-tap
 * (9,3,3,1)/16 upsample of its parent's quadrant, plus a byte residual, plus a noise layer that
 * modulates the colour directly or in HSV space
 * (terrain/sources/proland/ortho/OrthoProducer.cpp:268-372 + demo/shaders/ortho/upsampleOrthoShader.glsl).
 * Tiles are tile_w x tile_w RGBA8 with a 2-texel border (tile_w = tileSize + 4, e.g. 196);
 * (tile_w - 4) must be a multiple of 8 (196 and 100 in every shipped archive). */

/* createOrthoNoise (OrthoProducer.cpp:48-118): six tile_w x tile_w RGBA8 layers from the reference's
 * LCG streams, stored on the device in their 4 rotations.  host_out (optional) receives the six
 * unrotated layers, 6*tile_w*tile_w*4 bytes. */
int pl_ortho_noise_init(pl_ctx *ctx, int tile_w, uint8_t *host_out);
/* the six unrotated layers only (host code, no device needed) */
int pl_ortho_noise_host(int tile_w, uint8_t *out);

/* per-producer constants: the orthoProducer resource (OrthoProducer.cpp:440-560) */
typedef struct pl_ortho_scene {
    int32_t tile_w;            /* tileWidth uniform                                          */
    int32_t channels;          /* channels of the residual tiles (1..4); a missing channel reads
                                  0, a missing alpha 255 (GL texture swizzle defaults)       */
    int32_t hsv;               /* hsv="true": noiseUVLH.w                                    */
    int32_t face;              /* face="..." or the name's last digit                        */
    float scale;               /* scale="..." (default 2): residualOSH.w                     */
    float noise_color[4];      /* cnoise="..." / 255                                         */
    float root_noise_color[4]; /* rnoise="..." / 255                                         */
    int32_t n_amp;             /* noise="..." list (<= 32 levels)                            */
    int32_t max_level;         /* maxLevel (-1: none): hasTile                               */
    int32_t out_channels;      /* channels of the ortho storage: 0 or 4 = RGBA8; 1..3 (RGB8, RG8,
                                  R8): the alpha channel is not computed and its byte written as 0  */
    float noise_amp[32];
} pl_ortho_scene;

/* per-tile uniforms, OrthoProducer.cpp:286-366 */
typedef struct pl_ortho_req {
    int32_t out_slot;       /* GPUSlot::l of the tile being produced                         */
    int32_t parent_slot;    /* coarseLevelOSL.w, -1 at level 0                               */
    int32_t resid_slot;     /* slot in the residual pool, -1: residualOSH = -1               */
    int32_t dx, dy;         /* coarseLevelOSL.xy in texels: (t%2)*(tile_w-4)/2               */
    int32_t noise_r;        /* noiseUVLH.x (.y = (noise_r+1)%4)                              */
    int32_t noise_l;        /* noiseUVLH.z                                                   */
    int32_t level;
    float noise_color[4];   /* the noiseColor uniform (scaled by the level's amplitude)      */
    int32_t tx, ty;
    int32_t pad_[2];
} pl_ortho_req;             /* 64 bytes */

/* Fill one request exactly as OrthoProducer::doCreateTile does (host); slots are set to -1. */
void pl_ortho_make_req(const pl_ortho_scene *scene, int level, int tx, int ty, int has_resid,
                       pl_ortho_req *req);
/* n tiles of one level, consecutive in Morton order from morton0, into slots out_slot0.. (parents as in
 * pl_make_requests_range); nthreads host threads (<= 0: all).  No residuals. */
int pl_ortho_make_requests_range(const pl_ortho_scene *scene, int level, uint64_t morton0, int n,
                                 int out_slot0, int parent_slot0, uint64_t parent_morton0,
                                 pl_ortho_req *reqs, int nthreads);

/* The batched upsampleOrthoShader: n tiles, one CTA per tile; the parent quadrant (+ border) is staged
 * in shared memory by bulk copies (cp.async.bulk, mbarrier-signalled).  reqs is HOST memory.
 * ortho and resid are PL_POOL_ORTHO_UN8x4 pools of the same tile_w; resid may be NULL. */
int pl_ortho_batch(pl_ctx *ctx, const pl_ortho_scene *scene, pl_pool *ortho, pl_pool *resid, int n,
                   const pl_ortho_req *reqs);
int pl_ortho_batch_dev(pl_ctx *ctx, const pl_ortho_scene *scene, pl_pool *ortho, pl_pool *resid, int n,
                       const pl_ortho_req *dev_reqs);

/* pl_produce_range for ortho tiles: the requests of a Morton range are generated on the device (no per-tile data
 * crosses the PCIe bus), then the ortho kernel runs on them.  No residuals.  The generated requests are the ones
 * pl_ortho_make_requests_range builds on the host, byte for byte (pl_debug_download_requests returns them in its
 * elevation-request array: both request types are 64 bytes). */
int pl_ortho_produce_range(pl_ctx *ctx, const pl_ortho_scene *scene, pl_pool *ortho, int level, uint64_t morton0, int n,
                           int out_slot0, int parent_slot0, uint64_t parent_morton0);

/* OrthoCPUProducer::doCreateTile, the TIFF branch (ortho/OrthoCPUProducer.cpp:205-232): n blobs of an ortho
 * residual file (written by ColorMipmap::produceTile, preprocess/terrain/ColorMipmap.cpp:296-325: one
 * little-endian TIFF, one DEFLATE strip of tile_w * tile_w * channels bytes, channels = 1..4 samples of 8 bits),
 * inflated on the device into slots out_slots[j] of an RGBA8 pool (channels the file does not have are written as
 * 0).  blobs / offsets / sizes are HOST memory, as for pl_residual_decode_batch; the file handling (header of 7
 * ints, offset table, tile id = tx + ty * 2^level + (4^level - 1) / 3, OrthoCPUProducer.cpp:84-118,243-246)
 * stays with the caller.  *channels (optional) receives the sample count of the blobs (all must agree): the value
 * for pl_ortho_scene.channels.  DXT-compressed files (flags & 1) are consumed by the GL texture unit in the
 * reference and are not supported: PL_ERR_CORRUPT. */
int pl_ortho_decode_batch(pl_ctx *ctx, pl_pool *out, int n, const uint8_t *blobs, const uint64_t *offsets,
                          const uint32_t *sizes, const int32_t *out_slots, int *channels);

#ifdef __cplusplus
}
#endif
#endif
