/*
 * NormalProducer -- makes terrain normal tiles on the device from elevation tiles.
 *
 * Host mirror of terrain/sources/proland/dem/NormalProducer.h:54-178 /
 * NormalProducer.cpp:70-308: same overrides and dependencies (the elevation
 * tile of the same quad, and the parent normal tile), the uniforms of
 * NormalProducer.cpp:196-283 become one pl_norm_req per tile, launched in batches
 * through pl_normal_batch.  The storage format picks the output (RG8: fine normal;
 * RGBA8: fine + parent coarse normal); the elevation storage's filter is passed on
 * because the shader's +0.25 texel fetches see it.
 */
#ifndef PROLAND_B200_NORMAL_PRODUCER_H
#define PROLAND_B200_NORMAL_PRODUCER_H

#include <vector>

#include "proland/producer/GPUTileStorage.h"
#include "proland/producer/TileProducer.h"

namespace proland
{

PROLAND_API class NormalProducer : public TileProducer, public BatchSource
{
public:
    NormalProducer(ptr<TileCache> cache, ptr<TileProducer> elevationTiles, int gridSize, bool deform);
    virtual ~NormalProducer();

    virtual void getReferencedProducers(std::vector<ptr<TileProducer> > &producers) const;
    virtual void setRootQuadSize(float size);
    virtual int getBorder();
    virtual bool hasTile(int level, int tx, int ty);

    virtual void flushBatch();
    unsigned long getTileCount() const { return tileCount; }
    unsigned long getBatchCount() const { return batchCount; }

protected:
    NormalProducer();
    void init(ptr<TileCache> cache, ptr<TileProducer> elevationTiles, int gridSize, bool deform);

    virtual void *getContext() const;
    virtual ptr<Task> startCreateTile(int level, int tx, int ty, unsigned int deadline, ptr<Task> task,
                                      ptr<TaskGraph> owner);
    virtual void beginCreateTile();
    virtual bool doCreateTile(int level, int tx, int ty, TileStorage::Slot *data);
    virtual void endCreateTile();
    virtual void stopCreateTile(int level, int tx, int ty);

private:
    ptr<TileProducer> elevationTiles;
    bool deform;
    int gridMeshSize;
    ptr<DeviceContext> context;
    GPUTileStorage *storage;
    GPUTileStorage *elevationStorage;
    std::vector<pl_norm_req> pending;
    unsigned long tileCount, batchCount;

    pl_norm_scene scene() const;
};

}  // namespace proland

#endif
