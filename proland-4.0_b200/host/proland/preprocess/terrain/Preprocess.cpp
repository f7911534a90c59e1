#include "proland/preprocess/terrain/Preprocess.h"

#include <cassert>
#include <cerrno>
#include <cstdint>
#include <cstring>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include "proland/producer/DeviceContext.h"

namespace proland
{

InputMap::InputMap(int width, int height, int channels, int tileSize, int) :
    width(width), height(height), channels(channels), tileSize(tileSize)
{
    assert(tileSize > 0);
    assert(width % tileSize == 0);
    assert(height % tileSize == 0);
}

InputMap::~InputMap()
{
}

float *InputMap::getValues(int x, int y)
{
    float *v = new float[(size_t) tileSize * tileSize * channels];
    for (int j = 0; j < tileSize; ++j) {
        for (int i = 0; i < tileSize; ++i) {
            const vec4f c = getValue(x + i, y + j);
            const float in[4] = { c.x, c.y, c.z, c.w };
            memcpy(v + ((size_t) i + (size_t) j * tileSize) * channels, in, sizeof(float) * channels);
        }
    }
    return v;
}

vec4f InputMap::get(int x, int y)
{
    return getValue(max(min(x, width - 1), 0), max(min(y, height - 1), 0));
}

int ArrayInputMap::tileFor(int width, int height)
{
    int tile = 256;
    while (tile > 1 && (width % tile != 0 || height % tile != 0)) tile /= 2;
    return tile;
}

ArrayInputMap::ArrayInputMap(const float *data, int width, int height) :
    InputMap(width, height, 1, tileFor(width, height)), data(data)
{
}

vec4f ArrayInputMap::getValue(int x, int y)
{
    return vec4f(data[(size_t) y * width + x], 0, 0, 0);
}

float *ArrayInputMap::getValues(int x, int y)
{
    /* rows of the array are rows of the tile: no per-pixel virtual call */
    float *v = new float[(size_t) tileSize * tileSize];
    for (int j = 0; j < tileSize; ++j) {
        memcpy(v + (size_t) j * tileSize, data + (size_t) (y + j) * width + x, sizeof(float) * tileSize);
    }
    return v;
}

namespace
{

/* PLH_PREPROCESS_TIMING=1: where the builder's wall time goes, on stderr */
struct Phase
{
    const char *name;
    std::chrono::steady_clock::time_point t0;
    explicit Phase(const char *name) : name(name), t0(std::chrono::steady_clock::now()) {}
    ~Phase()
    {
        static const bool on = getenv("PLH_PREPROCESS_TIMING") != NULL;
        if (on) fprintf(stderr, "[preprocess] %-28s %8.1f ms\n", name, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }
};

bool fexists(const string &name)
{
    const int fd = open(name.c_str(), O_RDONLY);
    if (fd != -1) {
        close(fd);
        return true;
    }
    return false;
}

void createDir(const string &dir)
{
    if (mkdir(dir.c_str(), 0777) != 0 && errno != EEXIST) {
        if (Logger::ERROR_LOGGER != NULL) {
            Logger::ERROR_LOGGER->log("PREPROCESS", "Cannot create directory " + dir);
        }
        throw exception();
    }
}

/* the x channel of the whole map, row-major, read tile by tile (getValues) */
vector<float> readMap(InputMap *src)
{
    vector<float> m((size_t) src->width * src->height);
    const int ts = src->tileSize, ch = src->channels;
    for (int ty = 0; ty < src->height / ts; ++ty) {
        for (int tx = 0; tx < src->width / ts; ++tx) {
            float *v = src->getValues(tx * ts, ty * ts);
            for (int j = 0; j < ts; ++j) {
                for (int i = 0; i < ts; ++i) {
                    m[(size_t) (ty * ts + j) * src->width + tx * ts + i] = v[((size_t) i + (size_t) j * ts) * ch];
                }
            }
            delete[] v;
        }
    }
    return m;
}

struct Pools
{
    pl_pool *heights, *approx, *resid;
    Pools() : heights(NULL), approx(NULL), resid(NULL) {}
    ~Pools()
    {
        if (heights) pl_pool_destroy(heights);
        if (approx) pl_pool_destroy(approx);
        if (resid) pl_pool_destroy(resid);
    }
};

struct CubeHolder
{
    pl_height_cube *cube;
    CubeHolder() : cube(NULL) {}
    ~CubeHolder() { if (cube) pl_height_cube_destroy(cube); }
};

/* HeightMipmap::generate for one face (HeightMipmap.cpp:99-130, 255-324): level by level, every tile of the level in
 * one batch */
void generateFace(pl_ctx *ctx, pl_height_cube *cube, Pools &pools, int face, int topLevelSize, int baseLevelSize, int tileSize,
                  float scale, const string &file)
{
    int minLevel = 0, maxLevel = 0;      /* HeightMipmap.cpp:43-54 */
    for (int size = tileSize; size > topLevelSize; size /= 2) ++minLevel;
    for (int size = baseLevelSize; size > topLevelSize; size /= 2) ++maxLevel;
    const int n = tileSize + 5;
    const int nTilesTotal = minLevel + ((1 << (max(maxLevel - minLevel, 0) * 2 + 2)) - 1) / 3;
    vector<int16_t> tiles;
    vector<uint64_t> offsets((size_t) nTilesTotal + 1, 0);
    vector<int16_t> level_tiles;      /* the residual tiles of a level, n x n each, read back in one copy */
    vector<pl_height_req> hreqs;
    vector<pl_resid_enc_req> ereqs;
    size_t tileId = 0;
    for (int level = 0; level <= maxLevel; ++level) {
        const int nt = max(1, (baseLevelSize / tileSize) >> (maxLevel - level));
        const int ts = min(topLevelSize << level, tileSize);
        const int count = nt * nt;
        hreqs.assign(count, pl_height_req());
        ereqs.assign(count, pl_resid_enc_req());
        for (int ty = 0; ty < nt; ++ty) {
            for (int tx = 0; tx < nt; ++tx) {
                const int k = tx + ty * nt;
                pl_height_req &h = hreqs[k];
                h.face = face; h.level = level; h.tx = tx; h.ty = ty; h.out_slot = k;
                pl_resid_enc_req &e = ereqs[k];
                e.tile_slot = k;
                const int pnt = max(1, nt / 2);
                e.parent_slot = level == 0 ? -1 : (tx / 2) + (ty / 2) * pnt;
                e.approx_slot = k;
                e.resid_slot = k;
                e.tile_size = ts;
                e.tx = tx; e.ty = ty;
            }
        }
        DeviceContext::check(pl_height_tiles(ctx, cube, pools.heights, topLevelSize, tileSize, scale, count, hreqs.data()));
        /* one approximation pool, two halves: even levels write the lower half and read their parents from the
         * upper one, odd levels the other way round */
        const int half = pl_pool_capacity(pools.approx) / 2;
        for (int k = 0; k < count; ++k) {
            if (level > 0) ereqs[k].parent_slot += (level & 1) ? 0 : half;
            ereqs[k].approx_slot += (level & 1) ? half : 0;
        }
        DeviceContext::check(pl_residual_encode_batch(ctx, pools.heights, pools.approx, pools.resid, count, ereqs.data(), NULL, NULL));
        /* tiles in id order: levels below minLevel first, then row-major per level (ResidualProducer::getTileId) */
        level_tiles.resize((size_t) count * n * n);
        DeviceContext::check(pl_pool_download_range(pools.resid, 0, count, level_tiles.data(), sizeof(int16_t) * level_tiles.size()));
        const int w = ts + 5;
        tiles.reserve(tiles.size() + (size_t) count * w * w);
        for (int k = 0; k < count; ++k) {
            const int16_t *slot = level_tiles.data() + (size_t) k * n * n;
            offsets[tileId] = tiles.size();
            for (int j = 0; j < w; ++j) tiles.insert(tiles.end(), slot + (size_t) j * n, slot + (size_t) j * n + w);
            ++tileId;
        }
    }
    offsets[tileId] = tiles.size();
    assert((int) tileId == nTilesTotal);
    Phase pw("  of which: write the file");
    DeviceContext::check(pl_residual_write_file(file.c_str(), minLevel, maxLevel, tileSize, 0, 0, 0, scale, tiles.data(), offsets.data(), -1));
}

void preprocess(InputMap *src, int dstMinTileSize, int dstTileSize, int dstMaxLevel, const string &dstFolder, float residualScale, int nfaces)
{
    assert(dstTileSize % dstMinTileSize == 0);
    const int dstSize = dstTileSize << dstMaxLevel;
    ptr<DeviceContext> dc = DeviceContext::get();
    dc->flush();
    pl_ctx *ctx = dc->handle();
    vector<float> map;
    {
        Phase ph("read the source map");
        map = readMap(src);
    }
    CubeHolder holder;
    Phase *ph = new Phase("base level grids");
    if (nfaces == 6) {
        DeviceContext::check(pl_height_cube_from_latlon(ctx, dstSize, map.data(), src->width, src->height, &holder.cube));
    } else {
        DeviceContext::check(pl_height_cube_from_plane(ctx, dstSize, map.data(), src->width, src->height, &holder.cube));
    }
    delete ph;
    const int last = (dstSize / dstTileSize) * (dstSize / dstTileSize);
    Pools pools;
    DeviceContext::check(pl_pool_create(ctx, PL_POOL_RESID_F32, dstTileSize + 5, last, &pools.heights));
    DeviceContext::check(pl_pool_create(ctx, PL_POOL_RESID_F32, dstTileSize + 5, 2 * last, &pools.approx));
    DeviceContext::check(pl_pool_create(ctx, PL_POOL_RESID_I16, dstTileSize + 5, last, &pools.resid));
    for (int f = 0; f < nfaces; ++f) {
        string file = dstFolder + "/DEM";
        if (nfaces == 6) file += char('1' + f);
        file += ".dat";
        Phase pf("one face: tiles + file");
        generateFace(ctx, holder.cube, pools, f, dstMinTileSize, dstSize, dstTileSize, residualScale, file);
    }
    DeviceContext::check(pl_sync(ctx));
}

}

void preprocessDem(InputMap *src, int dstMinTileSize, int dstTileSize, int dstMaxLevel,
    const string &dstFolder, const string &, float residualScale)
{
    if (fexists(dstFolder + "/DEM.dat")) {
        return;
    }
    createDir(dstFolder);
    preprocess(src, dstMinTileSize, dstTileSize, dstMaxLevel, dstFolder, residualScale, 1);
}

void preprocessSphericalDem(InputMap *src, int dstMinTileSize, int dstTileSize, int dstMaxLevel,
    const string &dstFolder, const string &, float residualScale)
{
    if (fexists(dstFolder + "/DEM1.dat") && fexists(dstFolder + "/DEM2.dat") && fexists(dstFolder + "/DEM3.dat") &&
        fexists(dstFolder + "/DEM4.dat") && fexists(dstFolder + "/DEM5.dat") && fexists(dstFolder + "/DEM6.dat"))
    {
        return;
    }
    createDir(dstFolder);
    preprocess(src, dstMinTileSize, dstTileSize, dstMaxLevel, dstFolder, residualScale, 6);
}

}
