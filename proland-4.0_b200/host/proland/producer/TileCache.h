/*
 * TileCache -- which tiles are in a TileStorage, who uses them, and which unused
 * ones can be recycled (LRU).
 *
 * Same contract as the reference (producer/TileCache.h:60-425,
 * TileCache.cpp:150-336): used tiles (users > 0), unused tiles kept in LRU order
 * for reuse, and the creation tasks of evicted tiles remembered so that a tile
 * asked for again reuses its task.  getTile: hit in used -> users+1; hit in
 * unused -> back to used; miss -> free slot, else evict the least recently used
 * unused tile, else NULL (cache full).  putTile: users-1, at 0 the tile becomes the
 * most recently used unused tile.  prefetchTile: like a miss, but the tile starts
 * unused and its task gets deadline 1u<<31.
 */
#ifndef PROLAND_B200_TILE_CACHE_H
#define PROLAND_B200_TILE_CACHE_H

#include <list>
#include <map>
#include <mutex>
#include <unordered_map>
#include <stdexcept>
#include <string>
#include <vector>

#include "ork/ork_lite.h"
#include "proland/producer/TileStorage.h"

using namespace ork;

namespace proland
{

class TileProducer;

/* getTile found no free slot and no unused tile to evict while acquiring the inputs of a tile */
class CacheFullError : public std::runtime_error
{
public:
    CacheFullError(const std::string &what) : std::runtime_error(what) {}
};

PROLAND_API class TileCache : public Object
{
public:
    class Tile
    {
    public:
        typedef std::pair<int, std::pair<int, int> > Id;   /* (level, (tx, ty)) */
        typedef std::pair<int, Id> TId;                    /* (producer id, Id) */

        const int producerId;
        const int level;
        const int tx;
        const int ty;
        /* the task (graph) that produces this tile */
        const ptr<Task> task;

        Tile(int producerId, int level, int tx, int ty, ptr<Task> task, TileStorage::Slot *data);
        ~Tile();

        /* the slot, or NULL while the task is not done; check asserts done + slot id */
        TileStorage::Slot *getData(bool check = true);
        Id getId() const;
        TId getTId() const;
        static Id getId(int level, int tx, int ty);
        static TId getTId(int producerId, int level, int tx, int ty);

    private:
        TileStorage::Slot *data;
        int users;
        friend class TileCache;
        friend class CreateTile;
    };

    TileCache(ptr<TileStorage> storage, std::string name, ptr<Scheduler> scheduler = NULL);
    virtual ~TileCache();

    ptr<TileStorage> getStorage();
    ptr<Scheduler> getScheduler();
    const std::string &getName() const { return name; }
    int getUsedTiles();
    int getUnusedTiles();
    /* statistics the reference keeps but only logs (TileCache.cpp:236) */
    int getQueries() const { return queries; }
    int getMisses() const { return misses; }

    Tile *findTile(int producerId, int level, int tx, int ty, bool includeCache = false);
    Tile *getTile(int producerId, int level, int tx, int ty, unsigned int deadline, int *users = NULL);
    ptr<Task> prefetchTile(int producerId, int level, int tx, int ty);
    int putTile(Tile *t);
    void invalidateTiles(int producerId);
    void invalidateTile(int producerId, int level, int tx, int ty);

protected:
    TileCache();
    std::string name;
    void init(ptr<TileStorage> storage, std::string name, ptr<Scheduler> scheduler = NULL);

private:
    typedef std::list<Tile *> Order;
    /* the reference keeps its tiles in std::map keyed by the nested-pair TId (TileCache.h:271-289): every lookup is a
     * tree descent of four-int comparisons.  Same key, hashed: one 64-bit mix of (producer, level, tx, ty) */
    struct TIdHash
    {
        size_t operator()(const Tile::TId &id) const
        {
            unsigned long long h = ((unsigned long long) (unsigned int) id.first << 56) ^ ((unsigned long long) (unsigned int) id.second.first << 48) ^
                                   ((unsigned long long) (unsigned int) id.second.second.first << 24) ^ (unsigned int) id.second.second.second;
            h ^= h >> 29;
            h *= 0x9E3779B97F4A7C15ull;
            return (size_t) (h ^ (h >> 32));
        }
    };
    template <typename V> struct TileMap { typedef std::unordered_map<Tile::TId, V, TIdHash> type; };

    int nextProducerId;
    std::map<int, TileProducer *> producers;
    ptr<TileStorage> storage;
    ptr<Scheduler> scheduler;
    TileMap<Tile *>::type usedTiles;
    TileMap<Order::iterator>::type unusedTiles;
    Order unusedTilesOrder;               /* front = least recently used */
    TileMap<Task *>::type deletedTiles;
    int queries;
    int misses;
    std::recursive_mutex mutex;

    /* a slot for a new tile: a free one, else the LRU unused tile's (which is evicted) */
    TileStorage::Slot *acquireSlot();
    /* the creation task for tile `id` in slot `data`, reusing the task of an evicted incarnation */
    ptr<Task> makeTask(int producerId, const Tile::TId &id, int level, int tx, int ty, TileStorage::Slot *data,
                       unsigned int deadline, bool *reused);
    void rerun(ptr<Task> task, Task::reason r, unsigned int deadline);
    void createTileTaskDeleted(int producerId, int level, int tx, int ty);

    friend class TileProducer;

/* getTile found no free slot and no unused tile to evict while acquiring the inputs of a tile */
class CacheFullError : public std::runtime_error
{
public:
    CacheFullError(const std::string &what) : std::runtime_error(what) {}
};
    friend class CreateTile;
};

}  // namespace proland

#endif
