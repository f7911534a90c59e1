/*
 * pl_container.cu -- writer of the residual container ResidualProducer reads (host code only).
 *
 * Reference: HeightMipmap::generate (preprocess/terrain/HeightMipmap.cpp:99-130) writes the header
 * (minLevel, maxLevel, tileSize, rootLevel, rootTx, rootTy as int32, scale as float32), a table of
 * (begin, end) uint32 offsets per tile id relative to the end of the table, then one blob per tile:
 * the levels below minLevel first, then level by level in Lebesgue (Z) order
 * (produceTilesLebeguesOrder :645-655).  A blob is a little-endian TIFF with one DEFLATE strip of
 * 2 x 8-bit samples = little-endian int16 (produceTile :606-630, libtiff + zlib there, zlib here);
 * all-zero tiles share the blob of the first all-zero tile (:601-611, :636-638).  File format:
 * src/terrain/doc/overview.txt:147-216.  The compressed bytes depend on the zlib build, the decoded
 * residuals do not: tests read the file back with the oracle's reader and the device decoder.
 */
#include <zlib.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "pl_internal.h"

namespace {

void put16(std::vector<uint8_t> &b, unsigned v) { b.push_back((uint8_t) (v & 255)); b.push_back((uint8_t) (v >> 8)); }
void put32(std::vector<uint8_t> &b, uint32_t v) { put16(b, v & 0xffff); put16(b, v >> 16); }

/* one tile -> TIFF: header, strip at byte 8, IFD behind the (even-padded) strip */
int tiff_blob(const int16_t *tile, int w, int zlevel, std::vector<uint8_t> &out)
{
    const uLong raw_bytes = (uLong) w * w * 2;
    uLongf cap = compressBound(raw_bytes);
    std::vector<uint8_t> strip(cap);
    /* int16 little-endian is the host byte order on every platform CUDA runs on */
    if (compress2(strip.data(), &cap, reinterpret_cast<const Bytef *>(tile), raw_bytes, zlevel) != Z_OK) return -1;
    strip.resize(cap);
    const uint32_t strip_len = (uint32_t) strip.size();
    if (strip.size() % 2) strip.push_back(0);
    out.clear();
    out.push_back('I'); out.push_back('I'); put16(out, 42);
    put32(out, 8 + (uint32_t) strip.size());
    out.insert(out.end(), strip.begin(), strip.end());
    struct Tag { unsigned tag, type, count; uint32_t value; };
    const Tag tags[] = { { 256, 4, 1, (uint32_t) w }, { 257, 4, 1, (uint32_t) w }, { 258, 3, 2, 8u | (8u << 16) },
                         { 259, 3, 1, 32946 }, { 262, 3, 1, 1 }, { 273, 4, 1, 8 }, { 274, 3, 1, 4 }, { 277, 3, 1, 2 },
                         { 279, 4, 1, strip_len }, { 284, 3, 1, 1 } };
    put16(out, sizeof(tags) / sizeof(tags[0]));
    for (const Tag &t : tags) { put16(out, t.tag); put16(out, t.type); put32(out, t.count); put32(out, t.value); }
    put32(out, 0);
    return 0;
}

struct Writer {
    FILE *f;
    int min_level;
    const std::vector<std::vector<uint8_t> > *blobs;   /* compressed up front, on all host threads; empty = all-zero tile */
    const std::vector<char> *is_constant;
    std::vector<uint32_t> offsets;
    uint32_t offset;
    int constant_tile;
    int error;

    void produce(int level, int tx, int ty)
    {
        if (error) return;
        int tileid;
        if (level < min_level) {
            tileid = level;
        } else {
            const int l = level - min_level;
            tileid = min_level + tx + ty * (1 << l) + ((1 << (2 * l)) - 1) / 3;
        }
        const bool constant = (*is_constant)[tileid] != 0;
        if (constant && constant_tile != -1) {
            offsets[2 * tileid] = offsets[2 * constant_tile];
            offsets[2 * tileid + 1] = offsets[2 * constant_tile + 1];
        } else {
            const std::vector<uint8_t> &blob = (*blobs)[tileid];
            if (blob.empty() || fwrite(blob.data(), 1, blob.size(), f) != blob.size()) {
                error = 1;
                return;
            }
            offsets[2 * tileid] = offset;
            offset += (uint32_t) blob.size();
            offsets[2 * tileid + 1] = offset;
        }
        if (constant && constant_tile == -1) constant_tile = tileid;
    }

    /* produceTilesLebeguesOrder: the tiles of stored level min_level + l in Z order */
    void lebesgue(int l, int level, int tx, int ty)
    {
        if (level < l) {
            lebesgue(l, level + 1, 2 * tx, 2 * ty);
            lebesgue(l, level + 1, 2 * tx + 1, 2 * ty);
            lebesgue(l, level + 1, 2 * tx, 2 * ty + 1);
            lebesgue(l, level + 1, 2 * tx + 1, 2 * ty + 1);
        } else {
            produce(min_level + level, tx, ty);
        }
    }
};

}  // namespace

extern "C" int pl_residual_write_file(const char *path, int min_level, int max_level, int tile_size, int root_level,
                                      int root_tx, int root_ty, float scale, const int16_t *tiles,
                                      const uint64_t *tile_offsets, int zlib_level)
{
    if (!path || !tiles || !tile_offsets) return pl_set_error(PL_ERR_ARG, "NULL argument");
    if (min_level < 0 || max_level < 0 || min_level > 16 || max_level - min_level > 12 || tile_size < 2 ||
        (tile_size >> min_level) < 1 || zlib_level < -1 || zlib_level > 9)
        return pl_set_error(PL_ERR_ARG, "bad container parameters");
    const int d = max_level - min_level > 0 ? max_level - min_level : 0;
    const int ntiles = min_level + ((1 << (d * 2 + 2)) - 1) / 3;          /* HeightMipmap.cpp:105 */
    /* the blobs are independent: compress them on all host threads (libtiff + zlib do one at a time), then
     * lay them out in the reference's order.  The first all-zero tile in WRITING order keeps its blob. */
    std::vector<std::vector<uint8_t> > blobs((size_t) ntiles);
    std::vector<char> is_constant((size_t) ntiles, 0);
    {
        auto width_of = [&](int tileid) {
            int level = tileid;
            if (tileid >= min_level) {
                level = min_level;
                int first = min_level;
                while (first + (1 << (2 * (level - min_level))) <= tileid) { first += 1 << (2 * (level - min_level)); ++level; }
            }
            return (level < min_level ? tile_size >> (min_level - level) : tile_size) + 5;
        };
        std::atomic<int> next(0), failed(0);
        auto work = [&]() {
            for (int id = next.fetch_add(1); id < ntiles; id = next.fetch_add(1)) {
                const int wd = width_of(id);
                const int16_t *t = tiles + tile_offsets[id];
                bool constant = true;
                for (int i = 0; i < wd * wd && constant; ++i) constant = t[i] == 0;
                is_constant[id] = constant ? 1 : 0;
                if (tiff_blob(t, wd, zlib_level, blobs[id]) != 0) failed.store(1);
            }
        };
        int nt = (int) std::thread::hardware_concurrency();
        if (const char *e = getenv("PL_HOST_THREADS")) { const int cap = atoi(e); if (cap > 0 && nt > cap) nt = cap; }
        nt = nt < 1 ? 1 : (nt > 16 ? 16 : nt);
        if (nt > ntiles) nt = ntiles;
        std::vector<std::thread> pool;
        for (int t = 1; t < nt; ++t) pool.emplace_back(work);
        work();
        for (auto &th : pool) th.join();
        if (failed.load()) return pl_set_error(PL_ERR_IO, "zlib failed while compressing a tile");
    }
    FILE *f = fopen(path, "wb");
    if (!f) return pl_set_error(PL_ERR_IO, "cannot create %s", path);
    Writer w;
    w.blobs = &blobs;
    w.is_constant = &is_constant;
    w.f = f;
    w.min_level = min_level;
    w.offsets.assign((size_t) ntiles * 2, 0u);
    w.offset = 0;
    w.constant_tile = -1;
    w.error = 0;
    const int32_t head[6] = { min_level, max_level, tile_size, root_level, root_tx, root_ty };
    bool ok = fwrite(head, sizeof(head), 1, f) == 1 && fwrite(&scale, sizeof(float), 1, f) == 1 &&
              fwrite(w.offsets.data(), sizeof(uint32_t), w.offsets.size(), f) == w.offsets.size();
    if (ok) {
        for (int l = 0; l < min_level && l <= max_level; ++l) w.produce(l, 0, 0);
        for (int l = min_level; l <= max_level; ++l) w.lebesgue(l - min_level, 0, 0, 0);
        ok = !w.error && fseek(f, (long) (sizeof(head) + sizeof(float)), SEEK_SET) == 0 &&
             fwrite(w.offsets.data(), sizeof(uint32_t), w.offsets.size(), f) == w.offsets.size();
    }
    if (fclose(f) != 0) ok = false;
    if (!ok) return pl_set_error(PL_ERR_IO, "writing %s failed", path);
    return PL_OK;
}
