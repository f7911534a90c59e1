/*
 * TileLayer -- interface only.  Layers (producer/TileLayer.h in the reference)
 * modify tiles after their producer made them (blendShader pass of
 * ElevationProducer.cpp:380-395); they need the graph plugin and are OUT OF SCOPE
 * of the tile-production hot path.  The type exists so that
 * TileProducer keeps the reference's layer-related signatures.
 */
#ifndef PROLAND_B200_TILE_LAYER_H
#define PROLAND_B200_TILE_LAYER_H

#include "proland/producer/TileCache.h"

namespace proland
{

PROLAND_API class TileLayer : public Object
{
public:
    TileLayer(const char *type, bool deform = false) : Object(type), deform(deform), enabled(true) {}
    virtual ~TileLayer() {}
    bool isDeformed() { return deform; }
    bool isEnabled() { return enabled; }
    void setIsEnabled(bool e) { enabled = e; }
    virtual void setCache(ptr<TileCache> cache, int producerId) { (void) cache; (void) producerId; }
    virtual void setTileSize(int tileSize, int tileBorder, float rootQuadSize) { (void) tileSize; (void) tileBorder; (void) rootQuadSize; }
    virtual void useTile(int level, int tx, int ty, unsigned int deadline) { (void) level; (void) tx; (void) ty; (void) deadline; }
    virtual void unuseTile(int level, int tx, int ty) { (void) level; (void) tx; (void) ty; }
    virtual void prefetchTile(int level, int tx, int ty) { (void) level; (void) tx; (void) ty; }
    virtual void startCreateTile(int level, int tx, int ty, unsigned int deadline, ptr<Task> task, ptr<TaskGraph> owner)
    { (void) level; (void) tx; (void) ty; (void) deadline; (void) task; (void) owner; }
    virtual void beginCreateTile() {}
    virtual bool doCreateTile(int level, int tx, int ty, TileStorage::Slot *data) = 0;
    virtual void endCreateTile() {}
    virtual void stopCreateTile(int level, int tx, int ty) { (void) level; (void) tx; (void) ty; }

private:
    bool deform;
    bool enabled;
};

}  // namespace proland

#endif
