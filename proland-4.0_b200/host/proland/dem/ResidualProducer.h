/*
 * ResidualProducer -- loads elevation residual tiles from a residual file into a
 * float tile storage, decoding on the device.
 *
 * Host mirror of terrain/sources/proland/dem/ResidualProducer.h:60-274 /
 * ResidualProducer.cpp:60-384: same file format (7-word header, 2*ntiles offset
 * table, one TIFF blob per tile), same tile id / tile size / hasTile arithmetic,
 * same root composition for delta > 0, same nesting of sub-producers.  Different
 * underneath: the file is mapped once instead of fopen/fseek/fread per tile, and
 * libtiff + zlib + the int16 -> float loop are the batched device kernels behind
 * pl_residual_decode_batch / pl_residual_upsample (include/proland_b200.h).  The
 * tile lands in a device float storage (CPUTileStorage<float> here), which the
 * elevation kernel reads in place.
 */
#ifndef PROLAND_B200_RESIDUAL_PRODUCER_H
#define PROLAND_B200_RESIDUAL_PRODUCER_H

#include <string>
#include <vector>

#include "proland/producer/CPUTileStorage.h"
#include "proland/producer/TileProducer.h"

namespace proland
{

PROLAND_API class ResidualProducer : public TileProducer, public BatchSource
{
public:
    ResidualProducer(ptr<TileCache> cache, const char *name, int deltaLevel = 0, float zscale = 1.0);
    virtual ~ResidualProducer();

    virtual int getBorder();
    int getMinLevel();
    int getDeltaLevel();
    int getMaxLevel() const { return maxLevel; }
    void addProducer(ptr<ResidualProducer> p);
    virtual bool hasTile(int level, int tx, int ty);

    /* file-level arithmetic (public in spirit: the preprocessing tools and tests use it) */
    int getTileSize(int level);
    int getTileId(int level, int tx, int ty);

    virtual void flushBatch();
    unsigned long getTileCount() const { return tileCount; }

protected:
    ResidualProducer();
    void init(ptr<TileCache> cache, const char *name, int deltaLevel = 0, float zscale = 1.0);
    virtual bool doCreateTile(int level, int tx, int ty, TileStorage::Slot *data);
    virtual void endCreateTile();

private:
    struct Job
    {
        int level, tx, ty;   /* in the file's own quadtree */
        int slot;
        bool root;           /* compose stored levels 0..deltaLevel (ResidualProducer.cpp:218-228) */
    };

    std::string name;
    int tileSize;
    int rootLevel;
    int deltaLevel;
    int rootTx;
    int rootTy;
    int minLevel;
    int maxLevel;
    float scale;
    unsigned int header;
    std::vector<unsigned int> offsets;
    std::vector<ptr<ResidualProducer> > producers;

    /* the mapped file */
    const unsigned char *fileData;
    size_t fileSize;

    ptr<DeviceContext> context;
    GPUTileStorage *storage;
    std::vector<Job> pending;
    unsigned long tileCount;

    /* offset and size of a tile's blob in the file; throws DeviceError(PL_ERR_CORRUPT) when out of range */
    void blobOf(int tileid, uint64_t *offset, uint32_t *size) const;
    void decodeOne(int level, int tx, int ty, int outSlot, int addSlot);
};

}  // namespace proland

#endif
