/*
 * OrthoProducer -- makes ortho (colour) tiles on the device by upsampling the parent tile, adding a byte
 * residual tile and a noise layer, directly or in HSV space.
 *
 * Host mirror of terrain/sources/proland/ortho/OrthoProducer.h / OrthoProducer.cpp:120-438: same
 * TileProducer overrides, same dependencies (parent tile, residual tile (level, tx, ty)), same uniforms --
 * one pl_ortho_req per tile, queued and launched in batches through pl_ortho_batch
 * (include/proland_b200.h) instead of one drawQuad + copyPixels per tile.  The GL objects of the
 * reference constructor (orthoTexture, residualTexture, the upsample program) are gone.  Residual tiles
 * live in a device byte storage (RGBA8 / RGB8 gpuTileStorage of the same tile size) instead of a
 * CPUTileStorage<unsigned char> that is uploaded per tile (OrthoProducer.cpp:296-318).  Layers are out of
 * scope.
 */
#ifndef PROLAND_B200_ORTHO_PRODUCER_H
#define PROLAND_B200_ORTHO_PRODUCER_H

#include <vector>

#include "proland/producer/GPUTileStorage.h"
#include "proland/producer/TileProducer.h"

namespace proland
{

PROLAND_API class OrthoProducer : public TileProducer, public BatchSource
{
public:
    /* rootNoiseColor / noiseColor: 4 floats each, already divided by 255 (OrthoProducer.cpp:462-481) */
    OrthoProducer(ptr<TileCache> cache, ptr<TileProducer> residualTiles, const float rootNoiseColor[4],
                  const float noiseColor[4], std::vector<float> &noiseAmp, bool noiseHsv = false, float scale = 2.0f,
                  int maxLevel = -1, int face = 1);
    virtual ~OrthoProducer();

    virtual void getReferencedProducers(std::vector<ptr<TileProducer> > &producers) const;
    virtual void setRootQuadSize(float size);
    virtual int getBorder();
    virtual bool hasTile(int level, int tx, int ty);
    virtual bool prefetchTile(int level, int tx, int ty);

    virtual void flushBatch();
    unsigned long getTileCount() const { return tileCount; }
    unsigned long getBatchCount() const { return batchCount; }

protected:
    ptr<TileProducer> residualTiles;
    int face;

    OrthoProducer();
    void init(ptr<TileCache> cache, ptr<TileProducer> residualTiles, const float rootNoiseColor[4],
              const float noiseColor[4], std::vector<float> &noiseAmp, bool noiseHsv, float scale, int maxLevel, int face);

    virtual void *getContext() const;
    virtual ptr<Task> startCreateTile(int level, int tx, int ty, unsigned int deadline, ptr<Task> task,
                                      ptr<TaskGraph> owner);
    virtual void beginCreateTile();
    virtual bool doCreateTile(int level, int tx, int ty, TileStorage::Slot *data);
    virtual void endCreateTile();
    virtual void stopCreateTile(int level, int tx, int ty);

private:
    pl_ortho_scene scene;
    int maxLevel;
    ptr<DeviceContext> context;
    GPUTileStorage *storage;
    std::vector<pl_ortho_req> pending;
    unsigned long tileCount, batchCount;
};

}  // namespace proland

#endif
