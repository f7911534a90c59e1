/*
 * TileProducer -- the plugin contract of the tile-production path.
 *
 * Same virtuals, names and calling order as the reference
 * (producer/TileProducer.h:57-469, TileProducer.cpp:199-217,318-353,454-543,
 * 709-791): TileCache::getTile -> createTile -> startCreateTile (acquire the
 * tiles this one is made from, build the task graph), then when the scheduler
 * runs the task: beginCreateTile / doCreateTile(level, tx, ty, slot) /
 * endCreateTile, and stopCreateTile when the task is done.  Subclasses written
 * against the reference's TileProducer override the same methods here.
 *
 * What is different below the contract: GPU producers do not draw in
 * doCreateTile, they queue the tile on their DeviceContext; the queue is
 * launched as one batch per producer when the scheduler finishes a wave of
 * independent tasks (BatchScheduler) or, under any other scheduler, at
 * endCreateTile.  Renderer-side members of the reference class
 * (getGpuTileCoords, updateTileMap, update(SceneManager)) are out of scope.
 */
#ifndef PROLAND_B200_TILE_PRODUCER_H
#define PROLAND_B200_TILE_PRODUCER_H

#include <mutex>
#include <unordered_set>
#include <vector>

#include "proland/producer/TileCache.h"
#include "proland/producer/TileLayer.h"

using namespace ork;

namespace proland
{

PROLAND_API class TileProducer : public Object
{
public:
    TileProducer(const char *type, const char *taskType, ptr<TileCache> cache, bool gpuProducer);
    virtual ~TileProducer();

    float getRootQuadSize();
    virtual void setRootQuadSize(float size);
    int getId();
    virtual ptr<TileCache> getCache();
    bool isGpuProducer();
    virtual int getBorder();
    virtual bool hasTile(int level, int tx, int ty);
    bool hasChildren(int level, int tx, int ty);
    virtual TileCache::Tile *findTile(int level, int tx, int ty, bool includeCache = false, bool done = false);
    virtual TileCache::Tile *getTile(int level, int tx, int ty, unsigned int deadline);
    virtual bool prefetchTile(int level, int tx, int ty);
    virtual void putTile(TileCache::Tile *t);
    virtual void invalidateTiles();
    virtual void invalidateTile(int level, int tx, int ty);
    virtual void getReferencedProducers(std::vector<ptr<TileProducer> > &producers) const;

    int getLayerCount() const;
    ptr<TileLayer> getLayer(int index) const;
    bool hasLayers() const;
    void addLayer(ptr<TileLayer> l);

    const char *getTaskType() const { return taskType; }

protected:
    TileProducer(const char *type, const char *taskType);
    void init(ptr<TileCache> cache, bool gpuProducer);

    virtual void *getContext() const;
    virtual ptr<Task> startCreateTile(int level, int tx, int ty, unsigned int deadline, ptr<Task> task,
                                      ptr<TaskGraph> owner);
    virtual void beginCreateTile();
    virtual bool doCreateTile(int level, int tx, int ty, TileStorage::Slot *data);
    virtual void endCreateTile();
    virtual void stopCreateTile(int level, int tx, int ty);

    void removeCreateTile(Task *t);
    ptr<TaskGraph> createTaskGraph(ptr<Task> task);
    /* A tile this one is made from could not be acquired: TileCache::getTile returned NULL.  The
     * reference asserts here and logs "Insufficient tile cache size" in TileSampler
     * (terrain/TileSampler.cpp:441-444); here the message is logged and CacheFullError thrown.
     * startCreateTile must have released what it had acquired before calling this. */
    static void cacheFull(const char *producerType);

    friend class CreateTile;
    friend class CreateTileTaskGraph;

private:
    std::vector<ptr<TileLayer> > layers;
    /* the live CreateTile tasks / task graphs of this producer.  The reference keeps a vector and removes a dying task by
     * find + erase (TileProducer.cpp:533-543): linear in the number of live tasks per eviction -- with 12 000 cached tiles
     * that was 85 % of getTile once the cache evicts.  A hash set: O(1) */
    std::unordered_set<Task *> tasks;
    const char *taskType;
    ptr<TileCache> cache;
    bool gpuProducer;
    int id;
    float rootQuadSize;
    std::mutex mutex;

    ptr<Task> createTile(int level, int tx, int ty, TileStorage::Slot *data, unsigned int deadline, ptr<Task> old);

    friend class TileCache;
};

}  // namespace proland

#endif
