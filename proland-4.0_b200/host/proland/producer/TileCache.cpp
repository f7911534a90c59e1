#include "proland/producer/TileCache.h"

#include <cassert>

#include "proland/producer/TileProducer.h"

namespace proland
{

TileCache::Tile::Tile(int producerId, int level, int tx, int ty, ptr<Task> task, TileStorage::Slot *data) :
    producerId(producerId), level(level), tx(tx), ty(ty), task(task), data(data), users(0)
{
    assert(data != NULL);
}

TileCache::Tile::~Tile()
{
}

TileStorage::Slot *TileCache::Tile::getData(bool check)
{
    const bool isDone = task->isDone();
    assert(isDone || !check);
    assert(getTId() == data->id || !check);
    (void) check;
    return isDone ? data : NULL;
}

TileCache::Tile::Id TileCache::Tile::getId() const
{
    return getId(level, tx, ty);
}

TileCache::Tile::TId TileCache::Tile::getTId() const
{
    return getTId(producerId, level, tx, ty);
}

TileCache::Tile::Id TileCache::Tile::getId(int level, int tx, int ty)
{
    return std::make_pair(level, std::make_pair(tx, ty));
}

TileCache::Tile::TId TileCache::Tile::getTId(int producerId, int level, int tx, int ty)
{
    return std::make_pair(producerId, getId(level, tx, ty));
}

TileCache::TileCache(ptr<TileStorage> storage, std::string name, ptr<Scheduler> scheduler) : Object("TileCache")
{
    init(storage, name, scheduler);
}

TileCache::TileCache() : Object("TileCache")
{
}

void TileCache::init(ptr<TileStorage> storage, std::string name, ptr<Scheduler> scheduler)
{
    this->nextProducerId = 0;
    this->storage = storage;
    this->scheduler = scheduler;
    this->queries = 0;
    this->misses = 0;
    this->name = name;
    if (storage != NULL) {
        usedTiles.reserve(2 * (size_t) storage->getCapacity());
        unusedTiles.reserve(2 * (size_t) storage->getCapacity());
    }
}

TileCache::~TileCache()
{
    /* Unused tiles go first, newest first: a tile whose task never ran still pins the tiles it is made
     * from, and deleting it (and with it its task) releases them into the unused list. */
    while (!unusedTilesOrder.empty()) {
        Tile *t = unusedTilesOrder.back();
        unusedTilesOrder.pop_back();
        unusedTiles.erase(t->getTId());
        storage->deleteSlot(t->data);
        delete t;
    }
    /* users release their tiles before they drop the cache: the reference asserts usedTiles is empty
     * here (TileCache.cpp:113-117).  Tiles pinned by tasks that something else keeps alive end up here
     * too; they are reported and freed with the cache. */
    if (!usedTiles.empty()) {
        if (Logger::WARNING_LOGGER != NULL) {
            Logger::WARNING_LOGGER->logf("CACHE", "%s: %d tiles still in use when the cache is deleted", name.c_str(),
                                         (int) usedTiles.size());
        }
        for (TileMap<Tile *>::type::iterator i = usedTiles.begin(); i != usedTiles.end(); ++i) {
            storage->deleteSlot(i->second->data);
            delete i->second;
        }
        usedTiles.clear();
    }
    deletedTiles.clear();
}

ptr<TileStorage> TileCache::getStorage()
{
    return storage;
}

ptr<Scheduler> TileCache::getScheduler()
{
    return scheduler;
}

int TileCache::getUsedTiles()
{
    return (int) usedTiles.size();
}

int TileCache::getUnusedTiles()
{
    return (int) unusedTiles.size();
}

TileCache::Tile *TileCache::findTile(int producerId, int level, int tx, int ty, bool includeCache)
{
    assert(producers.find(producerId) != producers.end());
    std::lock_guard<std::recursive_mutex> lock(mutex);
    const Tile::TId id = Tile::getTId(producerId, level, tx, ty);
    TileMap<Tile *>::type::iterator u = usedTiles.find(id);
    if (u != usedTiles.end()) {
        return u->second;
    }
    if (includeCache) {
        TileMap<Order::iterator>::type::iterator c = unusedTiles.find(id);
        if (c != unusedTiles.end()) {
            return *(c->second);
        }
    }
    return NULL;
}

TileStorage::Slot *TileCache::acquireSlot()
{
    TileStorage::Slot *data = storage->newSlot();
    if (data == NULL && !unusedTilesOrder.empty()) {
        /* evict the least recently used unused tile; remember its task (TileCache.cpp:190-199) */
        Tile *victim = unusedTilesOrder.front();
        data = victim->data;
        unusedTiles.erase(victim->getTId());
        unusedTilesOrder.pop_front();
        deletedTiles.insert(std::make_pair(victim->getTId(), victim->task.get()));
        delete victim;   /* may destroy the task, which then erases itself from deletedTiles */
    }
    return data;
}

ptr<Task> TileCache::makeTask(int producerId, const Tile::TId &id, int level, int tx, int ty, TileStorage::Slot *data,
                              unsigned int deadline, bool *reused)
{
    ptr<Task> task;
    TileMap<Task *>::type::iterator d = deletedTiles.find(id);
    *reused = d != deletedTiles.end();
    if (*reused) {
        task = d->second;
        deletedTiles.erase(d);
    }
    try {
        return producers[producerId]->createTile(level, tx, ty, data, deadline, task);
    } catch (...) {
        /* the tile's inputs could not be acquired: the slot goes back, the error to the caller */
        storage->deleteSlot(data);
        throw;
    }
}

void TileCache::rerun(ptr<Task> task, Task::reason r, unsigned int deadline)
{
    if (scheduler == NULL) {
        task->setIsDone(false, 0, r);
    } else {
        scheduler->reschedule(task, r, deadline);
    }
}

TileCache::Tile *TileCache::getTile(int producerId, int level, int tx, int ty, unsigned int deadline, int *users)
{
    assert(producers.find(producerId) != producers.end());
    std::lock_guard<std::recursive_mutex> lock(mutex);
    const Tile::TId id = Tile::getTId(producerId, level, tx, ty);
    Tile *t = NULL;
    TileMap<Tile *>::type::iterator u = usedTiles.find(id);
    if (u != usedTiles.end()) {
        t = u->second;
    } else {
        ++queries;
        bool reused = false;
        TileMap<Order::iterator>::type::iterator c = unusedTiles.find(id);
        if (c != unusedTiles.end()) {
            /* still in storage: just back to the used set */
            t = *(c->second);
            unusedTilesOrder.erase(c->second);
            unusedTiles.erase(c);
        } else {
            TileStorage::Slot *data = acquireSlot();
            if (data != NULL) {
                ++misses;
                ptr<Task> task = makeTask(producerId, id, level, tx, ty, data, deadline, &reused);
                t = new Tile(producerId, level, tx, ty, task, data);
            }
        }
        if (t != NULL) {
            usedTiles.insert(std::make_pair(id, t));
            if (reused) {
                /* the data is gone but the task survived: it must run again (TileCache.cpp:224-233) */
                try {
                    rerun(t->task, Task::DATA_NEEDED, deadline);
                } catch (...) {
                    /* restarting the task could not re-acquire its inputs (cache full): the tile stays in
                     * the cache, unused, its task not done; the error goes to the caller */
                    usedTiles.erase(id);
                    unusedTiles[id] = unusedTilesOrder.insert(unusedTilesOrder.end(), t);
                    throw;
                }
            }
        }
        if (Logger::DEBUG_LOGGER != NULL) {
            Logger::DEBUG_LOGGER->logf("CACHE", "%s: tiles: %d used, %d reusable, total %d", name.c_str(),
                                       (int) usedTiles.size(), (int) unusedTiles.size(), storage->getCapacity());
        }
    }
    if (t != NULL) {
        if (users != NULL) {
            *users = t->users;
        }
        t->users += 1;
    }
    return t;
}

ptr<Task> TileCache::prefetchTile(int producerId, int level, int tx, int ty)
{
    assert(producers.find(producerId) != producers.end());
    std::lock_guard<std::recursive_mutex> lock(mutex);
    const Tile::TId id = Tile::getTId(producerId, level, tx, ty);
    ptr<Task> task;
    if (usedTiles.find(id) != usedTiles.end() || unusedTiles.find(id) != unusedTiles.end()) {
        return task;
    }
    TileStorage::Slot *data = acquireSlot();
    if (data != NULL) {
        const unsigned int deadline = 1u << 31u;
        bool reused = false;
        task = makeTask(producerId, id, level, tx, ty, data, deadline, &reused);
        Tile *t = new Tile(producerId, level, tx, ty, task, data);
        unusedTiles[id] = unusedTilesOrder.insert(unusedTilesOrder.end(), t);
        if (reused) {
            rerun(task, Task::DATA_NEEDED, deadline);
        }
    }
    return task;
}

int TileCache::putTile(Tile *t)
{
    std::lock_guard<std::recursive_mutex> lock(mutex);
    t->users -= 1;
    if (t->users == 0) {
        const Tile::TId id = t->getTId();
        TileMap<Tile *>::type::iterator u = usedTiles.find(id);
        assert(u != usedTiles.end() && u->second == t);
        usedTiles.erase(u);
        assert(unusedTiles.find(id) == unusedTiles.end());
        unusedTiles[id] = unusedTilesOrder.insert(unusedTilesOrder.end(), t);
    }
    return t->users;
}

void TileCache::invalidateTiles(int producerId)
{
    std::lock_guard<std::recursive_mutex> lock(mutex);
    const unsigned int deadline = 1u << 31u;
    for (TileMap<Tile *>::type::iterator i = usedTiles.begin(); i != usedTiles.end(); ++i) {
        if (i->second->producerId == producerId) rerun(i->second->task, Task::DATA_CHANGED, deadline);
    }
    for (Order::iterator j = unusedTilesOrder.begin(); j != unusedTilesOrder.end(); ++j) {
        if ((*j)->producerId == producerId) rerun((*j)->task, Task::DATA_CHANGED, deadline);
    }
    for (TileMap<Task *>::type::iterator k = deletedTiles.begin(); k != deletedTiles.end(); ++k) {
        if (k->first.first == producerId) rerun(k->second, Task::DATA_CHANGED, deadline);
    }
}

void TileCache::invalidateTile(int producerId, int level, int tx, int ty)
{
    std::lock_guard<std::recursive_mutex> lock(mutex);
    const Tile::TId id = Tile::getTId(producerId, level, tx, ty);
    const unsigned int deadline = 1u << 31u;
    TileMap<Tile *>::type::iterator u = usedTiles.find(id);
    if (u != usedTiles.end()) rerun(u->second->task, Task::DATA_CHANGED, deadline);
    TileMap<Order::iterator>::type::iterator c = unusedTiles.find(id);
    if (c != unusedTiles.end()) rerun((*(c->second))->task, Task::DATA_CHANGED, deadline);
    TileMap<Task *>::type::iterator k = deletedTiles.find(id);
    if (k != deletedTiles.end()) rerun(k->second, Task::DATA_CHANGED, deadline);
}

void TileCache::createTileTaskDeleted(int producerId, int level, int tx, int ty)
{
    std::lock_guard<std::recursive_mutex> lock(mutex);
    /* the task of an evicted tile died: forget it (tasks of tiles still in the cache are not in the map) */
    deletedTiles.erase(Tile::getTId(producerId, level, tx, ty));
}

}  // namespace proland
