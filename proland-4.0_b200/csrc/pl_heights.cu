/*
 * pl_heights.cu -- the height pyramid of the residual builder on the device: what HeightMipmap does BEFORE
 * buildResiduals (the part round 1 left to the caller).
 *
 * Reference (terrain/sources/proland/preprocess/terrain/):
 *   Preprocess.cpp:155-213, 406-445, 512-585   preprocessSphericalDem: six cube projections, SphericalHeightFunction
 *                                              (lon / lat bilinear fetch from the source map), one HeightMipmap per face
 *   HeightMipmap.cpp:67-81     setCube         neighbours + rotations across the 12 cube edges
 *   HeightMipmap.cpp:149-254   buildBaseLevelTiles / buildMipmapLevel: level l = every second sample of level l + 1
 *   HeightMipmap.cpp:327-372   getTileHeight   corner samples collapse onto the corner, samples past an edge come from
 *                                              the neighbouring face; AbstractTileCache.cpp:73-92 clamps a flat DEM
 *   HeightMipmap.cpp:404-412   getTile         the (ts + 5)^2 height tile the residual of a tile is computed from
 *
 * The reference writes every mipmap level to TIFF files and reads them back tile by tile.  Those files are the base
 * level decimated (stored_l(x, y) = base(x << (L - l), y << (L - l))), and their stitched borders are never read
 * back -- so here the six BASE grids stay resident in HBM (int16, (B + 1)^2 each: 57 MB per face at B = 5 376) and a
 * height tile of any level is one gather kernel: index arithmetic (corner rule, edge rule with rotation, clamp),
 * then base[face'][y' << s][x' << s].  Its output feeds pl_residual_encode_batch directly.
 */
#include <cmath>
#include <vector>

#include "pl_internal.h"

struct pl_height_cube {
    pl_ctx *ctx;
    int nfaces, B;
    short *dev[6];
};

namespace {

__constant__ int kNeigh[6][4] = { { 4, 2, 1, 3 }, { 4, 2, 5, 0 }, { 1, 3, 5, 0 }, { 2, 4, 5, 0 }, { 3, 1, 5, 0 }, { 4, 2, 3, 1 } };
__constant__ int kRot[6][4] = { { 3, 1, 0, 2 }, { 0, 0, 0, 0 }, { 0, 0, 1, 3 }, { 0, 0, 2, 2 }, { 0, 0, 3, 1 }, { 1, 3, 2, 0 } };

struct CubeView { const short *face[6]; int nfaces, B; };

__device__ __forceinline__ float cube_height(const CubeView &c, int maxLevel, int level, int face, int x, int y)
{
    const int levelSize = 1 + (c.B >> (maxLevel - level));
    if (c.nfaces == 6) {
        for (int hop = 0; hop < 8; ++hop) {
            if (x <= 2 && y <= 2) { x = 0; y = 0; }
            else if (x > levelSize - 4 && y <= 2) { x = levelSize - 1; y = 0; }
            else if (x <= 2 && y > levelSize - 4) { x = 0; y = levelSize - 1; }
            else if (x > levelSize - 4 && y > levelSize - 4) { x = levelSize - 1; y = levelSize - 1; }
            int side = -1, ax = 0, ay = 0;
            if (x < 0) { side = 0; ax = levelSize - 1 + x; ay = y; }
            else if (x >= levelSize) { side = 1; ax = x - levelSize + 1; ay = y; }
            else if (y < 0) { side = 2; ax = x; ay = levelSize - 1 + y; }
            else if (y >= levelSize) { side = 3; ax = x; ay = y - levelSize + 1; }
            if (side < 0) break;
            const int r = kRot[face][side], n = levelSize;
            switch (r) {      /* rotation(), ColorMipmap.cpp:421-441 */
            case 0: x = ax; y = ay; break;
            case 1: x = ay; y = n - 1 - ax; break;
            case 2: x = n - 1 - ax; y = n - 1 - ay; break;
            default: x = n - 1 - ay; y = ax; break;
            }
            face = kNeigh[face][side];
        }
    }
    const int w = levelSize - 1;
    x = min(max(x, 0), w);
    y = min(max(y, 0), w);
    const int sh = maxLevel - level;
    return (float) c.face[face][((size_t) y << sh) * (size_t) (c.B + 1) + ((size_t) x << sh)];
}

__global__ void __launch_bounds__(256) height_tiles_kernel(const CubeView c, const pl_height_req *reqs, int maxLevel, int topLevelSize,
                                                           int tileSize, float scale, unsigned char *pool, size_t slot_bytes, int pitch)
{
    const pl_height_req q = reqs[blockIdx.x];
    int ts = topLevelSize << q.level;
    if (ts > tileSize) ts = tileSize;
    const int w = ts + 5;
    float *dst = reinterpret_cast<float *>(pool + (size_t) q.out_slot * slot_bytes);
    for (int k = threadIdx.x; k < w * w; k += blockDim.x) {
        const int j = k / w, i = k - j * w;
        dst[(size_t) j * pitch + i] = cube_height(c, maxLevel, q.level, q.face, i + ts * q.tx - 2, j + ts * q.ty - 2) / scale;
    }
}

/* the direction of base sample (x, y) of a face: projection1..6, Preprocess.cpp:155-213 */
__host__ __device__ inline void cube_projection(int face, int x, int y, int B, double &sx, double &sy, double &sz)
{
    const double xl = (double) x / B * 2.0 - 1.0, yl = (double) y / B * 2.0 - 1.0;
    const double l = sqrt(xl * xl + yl * yl + 1.0);
    switch (face) {
    case 0: sx = xl / l; sy = yl / l; sz = 1.0 / l; break;
    case 1: sx = xl / l; sy = -1.0 / l; sz = yl / l; break;
    case 2: sx = 1.0 / l; sy = xl / l; sz = yl / l; break;
    case 3: sx = -xl / l; sy = 1.0 / l; sz = yl / l; break;
    case 4: sx = -1.0 / l; sy = -xl / l; sz = yl / l; break;
    default: sx = xl / l; sy = -yl / l; sz = -1.0 / l; break;
    }
}

/* SphericalHeightFunction::getHeight(lon, lat) (Preprocess.cpp:429-444): bilinear fetch, double products and sums, one
 * rounding to float; *slack: a bound on |dh| for an error of kAngleSlack source texels in lon / lat */
#define PL_ANGLE_SLACK 1e-9
__host__ __device__ inline double latlon_height(const float *src, int sw, int sh, double lon, double lat, double *slack)
{
    lon = lon / M_PI * (sw / 2);
    lat = lat / M_PI * sh;
    const int ilon = (int) floor(lon), ilat = (int) floor(lat);
    lon -= ilon;
    lat -= ilat;
    const double clon = 1.0 - lon, clat = 1.0 - lat;
    const int r0 = ilat < 0 ? 0 : (ilat > sh - 1 ? sh - 1 : ilat), r1 = ilat + 1 < 0 ? 0 : (ilat + 1 > sh - 1 ? sh - 1 : ilat + 1);   /* InputMap::get clamps */
    const double h1 = src[(size_t) r0 * sw + (ilon + sw) % sw], h2 = src[(size_t) r0 * sw + (ilon + sw + 1) % sw];
    const double h3 = src[(size_t) r1 * sw + (ilon + sw) % sw], h4 = src[(size_t) r1 * sw + (ilon + sw + 1) % sw];
    if (slack) {
        /* a texel boundary within the slack: the cell itself is uncertain -- flag unconditionally */
        const bool edge = lon < PL_ANGLE_SLACK || clon < PL_ANGLE_SLACK || lat < PL_ANGLE_SLACK || clat < PL_ANGLE_SLACK;
        *slack = edge ? 1e30 : PL_ANGLE_SLACK * (fabs(h2 - h1) + fabs(h4 - h3) + fabs(h3 - h1) + fabs(h4 - h2)) + 1e-9;
    }
#ifdef __CUDA_ARCH__
    return __dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(h1, clon), __dmul_rn(h2, lon)), clat), __dmul_rn(__dadd_rn(__dmul_rn(h3, clon), __dmul_rn(h4, lon)), lat));
#else
    return (h1 * clon + h2 * lon) * clat + (h3 * clon + h4 * lon) * lat;      /* host TU: -ffp-contract=off */
#endif
}

/* SphericalHeightFunction::getHeight + (short) h of buildBaseLevelTile: the base grid of one face from the source map.
 * atan2 / acos of the device are not bit-identical to the host libm the reference runs on; a sample whose int16 value
 * could change under that difference (the truncation (short)(float) h decides differently at h - slack and h + slack) is
 * listed in `unsure` and recomputed by the caller with the host's libm: the grid is the reference's, sample for sample. */
__global__ void __launch_bounds__(256) spherical_base_kernel(const float *src, int sw, int sh, int face, int B, short *out,
                                                             unsigned long long *unsure, unsigned int *n_unsure, unsigned int cap)
{
    const size_t n = (size_t) (B + 1) * (B + 1);
    for (size_t k = blockIdx.x * (size_t) blockDim.x + threadIdx.x; k < n; k += (size_t) gridDim.x * blockDim.x) {
        const int y = (int) (k / (B + 1)), x = (int) (k - (size_t) y * (B + 1));
        double sx, sy, sz, slack;
        cube_projection(face, x, y, B, sx, sy, sz);
        const double h = latlon_height(src, sw, sh, atan2(sy, sx) + M_PI, acos(sz), &slack);
        out[k] = (short) (float) h;
        if ((short) (float) (h - slack) != (short) (float) (h + slack)) {
            const unsigned int slot = atomicAdd(n_unsure, 1u);
            if (slot < cap) unsure[slot] = k;
        }
    }
}

__global__ void patch_samples_kernel(const unsigned long long *patch, unsigned int n, short *out)
{
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[patch[i] & 0xFFFFFFFFFFFFull] = (short) (unsigned short) (patch[i] >> 48);
}

/* PlaneHeightFunction::getHeight (Preprocess.cpp:335-366) + (short) h: the base grid of a flat DEM from the source map */
__global__ void __launch_bounds__(256) plane_base_kernel(const float *src, int sw, int sh, int B, short *out)
{
    const size_t n = (size_t) (B + 1) * (B + 1);
    for (size_t k = blockIdx.x * (size_t) blockDim.x + threadIdx.x; k < n; k += (size_t) gridDim.x * blockDim.x) {
        const int gy = (int) (k / (B + 1)), gx = (int) (k - (size_t) gy * (B + 1));
        double x = __dmul_rn(__ddiv_rn((double) gx, (double) B), (double) sw), y = __dmul_rn(__ddiv_rn((double) gy, (double) B), (double) sh);
        const int ix = (int) floor(x), iy = (int) floor(y);
        x -= ix;
        y -= iy;
        const double cx = 1.0 - x, cy = 1.0 - y;
        const int c0 = min(max(ix, 0), sw - 1), c1 = min(max(ix + 1, 0), sw - 1), r0 = min(max(iy, 0), sh - 1), r1 = min(max(iy + 1, 0), sh - 1);
        const float h1 = src[(size_t) r0 * sw + c0], h2 = src[(size_t) r0 * sw + c1], h3 = src[(size_t) r1 * sw + c0], h4 = src[(size_t) r1 * sw + c1];
        const float h = (float) (__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn((double) h1, cx), __dmul_rn((double) h2, x)), cy),
                                          __dmul_rn(__dadd_rn(__dmul_rn((double) h3, cx), __dmul_rn((double) h4, x)), y)));
        out[k] = (short) h;
    }
}

}  // namespace

static void cube_free(pl_height_cube *c)
{
    for (int f = 0; f < 6; ++f)
        if (c->dev[f]) cudaFree(c->dev[f]);
    delete c;
}

extern "C" int pl_height_cube_create(pl_ctx *ctx, int base_size, int nfaces, const int16_t *const *faces, pl_height_cube **out)
{
    if (!ctx || !faces || !out || base_size < 1 || (nfaces != 1 && nfaces != 6)) return pl_set_error(PL_ERR_ARG, "bad argument");
    PL_CUDA(cudaSetDevice(ctx->device));
    pl_height_cube *c = new pl_height_cube();
    c->ctx = ctx;
    c->nfaces = nfaces;
    c->B = base_size;
    for (int f = 0; f < 6; ++f) c->dev[f] = nullptr;
    const size_t bytes = sizeof(short) * (size_t) (base_size + 1) * (base_size + 1);
    for (int f = 0; f < nfaces; ++f) {
        cudaError_t e = faces[f] ? cudaMalloc(&c->dev[f], bytes) : cudaErrorInvalidValue;
        if (e == cudaSuccess) e = cudaMemcpyAsync(c->dev[f], faces[f], bytes, cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) {
            cube_free(c);
            return pl_set_error(PL_ERR_CUDA, "pl_height_cube_create: %s", cudaGetErrorString(e));
        }
    }
    PL_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = c;
    return PL_OK;
}

/* one face of the spherical base level; the samples the kernel could not decide are redone with the host's libm */
static cudaError_t spherical_face(pl_ctx *ctx, const float *dsrc, const float *src, int sw, int sh, int face, int B, short *dst,
                                  unsigned long long **d_unsure, unsigned int **d_count, unsigned int *cap)
{
    cudaError_t e = cudaSuccess;
    if (!*d_count) e = cudaMalloc(d_count, sizeof(unsigned int));
    for (int attempt = 0; attempt < 8 && e == cudaSuccess; ++attempt) {
        if (!*d_unsure) e = cudaMalloc(d_unsure, sizeof(unsigned long long) * (size_t) *cap);
        if (e == cudaSuccess) e = cudaMemsetAsync(*d_count, 0, sizeof(unsigned int), ctx->stream);
        if (e != cudaSuccess) break;
        spherical_base_kernel<<<4 * ctx->sm_count, 256, 0, ctx->stream>>>(dsrc, sw, sh, face, B, dst, *d_unsure, *d_count, *cap);
        unsigned int count = 0;
        e = cudaMemcpyAsync(&count, *d_count, sizeof(count), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) break;
        if (count > *cap) {      /* the list overflowed: a larger one, again */
            cudaFree(*d_unsure);
            *d_unsure = nullptr;
            *cap = count + (count >> 2);
            continue;
        }
        ctx->height_unsure += count;
        if (count == 0) return cudaSuccess;
        std::vector<unsigned long long> idx(count);
        e = cudaMemcpy(idx.data(), *d_unsure, sizeof(unsigned long long) * count, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) break;
        /* the value rides in the top 16 bits of the index word (an index is below 2^48) */
        for (unsigned int i = 0; i < count; ++i) {
            const int y = (int) (idx[i] / (unsigned long long) (B + 1)), x = (int) (idx[i] - (unsigned long long) y * (B + 1));
            double sx, sy, sz;
            cube_projection(face, x, y, B, sx, sy, sz);
            const short v = (short) (float) latlon_height(src, sw, sh, atan2(sy, sx) + M_PI, acos(sz), nullptr);
            idx[i] |= (unsigned long long) (unsigned short) v << 48;
        }
        e = cudaMemcpy(*d_unsure, idx.data(), sizeof(unsigned long long) * count, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) break;
        patch_samples_kernel<<<(count + 255) / 256, 256, 0, ctx->stream>>>(*d_unsure, count, dst);
        ctx->launches += 1;
        return cudaGetLastError();
    }
    return e == cudaSuccess ? cudaErrorUnknown : e;
}

static int cube_from_map(pl_ctx *ctx, int base_size, const float *src, int src_w, int src_h, int nfaces, pl_height_cube **out)
{
    if (!ctx || !src || !out || base_size < 1 || src_w < 2 || src_h < 1) return pl_set_error(PL_ERR_ARG, "bad argument");
    PL_CUDA(cudaSetDevice(ctx->device));
    pl_height_cube *c = new pl_height_cube();
    c->ctx = ctx;
    c->nfaces = nfaces;
    c->B = base_size;
    for (int f = 0; f < 6; ++f) c->dev[f] = nullptr;
    float *dsrc = nullptr;
    unsigned long long *d_unsure = nullptr;
    unsigned int *d_count = nullptr;
    unsigned int cap = 1u << 16;
    const size_t sbytes = sizeof(float) * (size_t) src_w * src_h, bytes = sizeof(short) * (size_t) (base_size + 1) * (base_size + 1);
    cudaError_t e = cudaMalloc(&dsrc, sbytes);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dsrc, src, sbytes, cudaMemcpyHostToDevice, ctx->stream);
    for (int f = 0; f < nfaces && e == cudaSuccess; ++f) {
        e = cudaMalloc(&c->dev[f], bytes);
        if (e != cudaSuccess) break;
        if (nfaces == 6) e = spherical_face(ctx, dsrc, src, src_w, src_h, f, base_size, c->dev[f], &d_unsure, &d_count, &cap);
        else plane_base_kernel<<<4 * ctx->sm_count, 256, 0, ctx->stream>>>(dsrc, src_w, src_h, base_size, c->dev[f]);
        if (e == cudaSuccess) e = cudaGetLastError();
        ctx->launches += 1;
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (dsrc) cudaFree(dsrc);
    if (d_unsure) cudaFree(d_unsure);
    if (d_count) cudaFree(d_count);
    if (e != cudaSuccess) {
        cube_free(c);
        return pl_set_error(PL_ERR_CUDA, "pl_height_cube_from_map: %s", cudaGetErrorString(e));
    }
    *out = c;
    return PL_OK;
}

extern "C" int pl_height_cube_from_latlon(pl_ctx *ctx, int base_size, const float *src, int src_w, int src_h, pl_height_cube **out)
{
    return cube_from_map(ctx, base_size, src, src_w, src_h, 6, out);
}

extern "C" int pl_height_cube_from_plane(pl_ctx *ctx, int base_size, const float *src, int src_w, int src_h, pl_height_cube **out)
{
    return cube_from_map(ctx, base_size, src, src_w, src_h, 1, out);
}

extern "C" uint64_t pl_debug_height_unsure(const pl_ctx *ctx) { return ctx ? ctx->height_unsure : 0; }

extern "C" void pl_height_cube_destroy(pl_height_cube *c)
{
    if (!c) return;
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    cube_free(c);
}

extern "C" int pl_height_cube_download(pl_ctx *ctx, const pl_height_cube *c, int face, int16_t *out)
{
    if (!ctx || !c || !out || face < 0 || face >= c->nfaces) return pl_set_error(PL_ERR_ARG, "bad argument");
    PL_CUDA(cudaSetDevice(ctx->device));
    PL_CUDA(cudaMemcpyAsync(out, c->dev[face], sizeof(short) * (size_t) (c->B + 1) * (c->B + 1), cudaMemcpyDeviceToHost, ctx->stream));
    PL_CUDA(cudaStreamSynchronize(ctx->stream));
    return PL_OK;
}

extern "C" int pl_height_tiles(pl_ctx *ctx, const pl_height_cube *cube, pl_pool *heights, int top_level_size, int tile_size,
                               float scale, int n, const pl_height_req *reqs)
{
    if (!ctx || !cube || !heights || n < 0) return pl_set_error(PL_ERR_ARG, "bad argument");
    if (n == 0) return PL_OK;
    if (!reqs) return pl_set_error(PL_ERR_ARG, "reqs is NULL");
    if (heights->kind != PL_POOL_RESID_F32) return pl_set_error(PL_ERR_ARG, "heights must be an F32 residual pool");
    if (top_level_size < 1 || tile_size < top_level_size || tile_size + 5 > heights->tile_w || scale == 0.0f)
        return pl_set_error(PL_ERR_ARG, "tile size %d does not fit the pool", tile_size);
    /* minLevel / maxLevel as the HeightMipmap constructor derives them (HeightMipmap.cpp:43-54) */
    int maxLevel = 0;
    for (int size = cube->B; size > top_level_size; size /= 2) ++maxLevel;
    if ((top_level_size << maxLevel) != cube->B) return pl_set_error(PL_ERR_ARG, "base size %d is not top_level_size << maxLevel", cube->B);
    for (int i = 0; i < n; ++i) {
        const pl_height_req &q = reqs[i];
        int ts = top_level_size << (q.level < 0 ? 0 : (q.level > 30 ? 30 : q.level));
        if (ts > tile_size) ts = tile_size;
        const int nt = q.level < 0 || q.level > maxLevel ? 0 : ((cube->B >> (maxLevel - q.level)) / ts);
        if (q.face < 0 || q.face >= cube->nfaces || q.level < 0 || q.level > maxLevel || q.tx < 0 || q.ty < 0 || q.tx >= nt || q.ty >= nt ||
            q.out_slot < 0 || q.out_slot >= heights->capacity)
            return pl_set_error(PL_ERR_ARG, "request %d: (face %d, level %d, %d, %d) -> slot %d out of range", i, q.face, q.level, q.tx, q.ty, q.out_slot);
    }
    PL_CUDA(cudaSetDevice(ctx->device));
    void *dev = nullptr;
    int rc = pl_stage_requests(ctx, reqs, sizeof(pl_height_req) * (size_t) n, &dev);
    if (rc) return rc;
    CubeView v;
    for (int f = 0; f < 6; ++f) v.face[f] = cube->dev[f];
    v.nfaces = cube->nfaces;
    v.B = cube->B;
    pl_timing_begin(ctx, PL_K_RESIDUAL, n);
    height_tiles_kernel<<<n, 256, 0, ctx->stream>>>(v, static_cast<const pl_height_req *>(dev), maxLevel, top_level_size, tile_size, scale,
                                                   heights->base, heights->slot_bytes, heights->pitch);
    pl_timing_end(ctx);
    PL_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return PL_OK;
}
