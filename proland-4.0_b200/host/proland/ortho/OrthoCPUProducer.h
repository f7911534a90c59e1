/*
 * OrthoCPUProducer -- loads ortho residual byte tiles from a file into a byte tile storage, decoding on
 * the device.
 *
 * Host mirror of terrain/sources/proland/ortho/OrthoCPUProducer.h / OrthoCPUProducer.cpp:60-276: same file
 * format (7-int header, (begin, end) int64 offsets per tile, one TIFF blob per tile), same tile id and
 * hasTile arithmetic, same "no file: all-zero tiles" mode.  Different underneath: the file is mapped once
 * instead of fopen / fseek64 / fread per tile, libtiff + zlib are the batched device kernels behind
 * pl_ortho_decode_batch (include/proland_b200.h), and the tile lands in a device byte storage
 * (cpuByteTileStorage -> CPUTileStorage<unsigned char> here) that OrthoProducer reads in place instead of
 * uploading it per tile.  DXT files (flags & 1) are refused: their blobs are GL texture data.
 */
#ifndef PROLAND_B200_ORTHO_CPU_PRODUCER_H
#define PROLAND_B200_ORTHO_CPU_PRODUCER_H

#include <string>
#include <vector>

#include "proland/producer/CPUTileStorage.h"
#include "proland/producer/TileProducer.h"

namespace proland
{

PROLAND_API class OrthoCPUProducer : public TileProducer, public BatchSource
{
public:
    OrthoCPUProducer(ptr<TileCache> cache, const char *name);
    virtual ~OrthoCPUProducer();

    virtual int getBorder();
    virtual bool hasTile(int level, int tx, int ty);
    /* true for DXT files (never here: they are refused at load time) */
    bool isCompressed();
    int getChannels() const { return channels; }
    int getMaxLevel() const { return maxLevel; }
    int getTileId(int level, int tx, int ty);

    virtual void flushBatch();
    unsigned long getTileCount() const { return tileCount; }

protected:
    OrthoCPUProducer();
    void init(ptr<TileCache> cache, const char *name);
    void load(const char *name);
    virtual bool doCreateTile(int level, int tx, int ty, TileStorage::Slot *data);
    virtual void endCreateTile();

private:
    struct Job { int tileid, slot; };

    std::string name;
    int channels;
    int tileSize;
    int border;
    int maxLevel;
    bool dxt;
    unsigned int header;
    std::vector<long long> offsets;
    const unsigned char *fileData;
    size_t fileSize;

    ptr<DeviceContext> context;
    GPUTileStorage *storage;
    std::vector<Job> pending;
    unsigned long tileCount;
};

}  // namespace proland

#endif
