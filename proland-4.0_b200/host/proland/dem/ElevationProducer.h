/*
 * ElevationProducer -- makes elevation tiles (zf, zc, zm) on the device by
 * upsampling the parent tile, adding amplitude-scaled noise and the residual tile.
 *
 * Host mirror of terrain/sources/proland/dem/ElevationProducer.h:57-248 /
 * ElevationProducer.cpp:157-455: same TileProducer overrides, same
 * dependencies (parent tile, residual tile (level, tx/mod, ty/mod)), same
 * uniforms -- which here become one pl_elev_req per tile, queued and launched in
 * batches through pl_elevation_batch (include/proland_b200.h) instead of one
 * drawQuad + copyPixels per tile.  The GL objects of the reference constructor
 * (textures, programs) are gone; the shader variant they selected is the
 * `Variant` argument.  Layers / blendShader are out of scope.
 */
#ifndef PROLAND_B200_ELEVATION_PRODUCER_H
#define PROLAND_B200_ELEVATION_PRODUCER_H

#include <vector>

#include "proland/producer/GPUTileStorage.h"
#include "proland/producer/TileProducer.h"

namespace proland
{

PROLAND_API class ElevationProducer : public TileProducer, public BatchSource
{
public:
    /* the upsampleShader flavours of the reference archives */
    struct Variant
    {
        bool slopeNoise;   /* demo / terrain2: noise amplitude modulated by slope and curvature */
        bool noClamp;      /* upsampleShader-noClamp.xml: zm = zf instead of max(zf, 0) */
        Variant(bool slopeNoise = true, bool noClamp = false) : slopeNoise(slopeNoise), noClamp(noClamp) {}
    };

    ElevationProducer(ptr<TileCache> cache, ptr<TileProducer> residualTiles, int gridMeshSize,
                      std::vector<float> &noiseAmp, bool flipDiagonals = false, int face = 0,
                      Variant variant = Variant());
    virtual ~ElevationProducer();

    virtual void getReferencedProducers(std::vector<ptr<TileProducer> > &producers) const;
    virtual void setRootQuadSize(float size);
    virtual int getBorder();

    /* per-tile (zmin, zmax) of zm over the tile interior, what TileSamplerZ reads back
     * (core/sources/proland/terrain/TileSamplerZ.cpp:253-351); the tile must be done */
    void getTileMinMax(TileCache::Tile *t, float *zmin, float *zmax);

    virtual void flushBatch();
    /* tiles produced so far (doCreateTile calls) and batches launched */
    unsigned long getTileCount() const { return tileCount; }
    unsigned long getBatchCount() const { return batchCount; }

protected:
    ptr<TileProducer> residualTiles;
    int face;

    ElevationProducer();
    void init(ptr<TileCache> cache, ptr<TileProducer> residualTiles, int gridMeshSize, std::vector<float> &noiseAmp,
              bool flipDiagonals, int face, Variant variant);

    virtual void *getContext() const;
    virtual ptr<Task> startCreateTile(int level, int tx, int ty, unsigned int deadline, ptr<Task> task,
                                      ptr<TaskGraph> owner);
    virtual void beginCreateTile();
    virtual bool doCreateTile(int level, int tx, int ty, TileStorage::Slot *data);
    virtual void endCreateTile();
    virtual void stopCreateTile(int level, int tx, int ty);

private:
    std::vector<float> noiseAmp;
    int gridMeshSize;
    bool flipDiagonals;
    Variant variant;
    ptr<DeviceContext> context;
    GPUTileStorage *storage;             /* the cache's storage */
    std::vector<pl_elev_req> pending;
    unsigned long tileCount, batchCount;

    int residualMod() const;
};

}  // namespace proland

#endif
