#include "proland/producer/TileProducer.h"

#include <algorithm>
#include <cassert>

namespace proland
{

/*
 * The task that makes one tile (TileProducer.cpp:44-247 of the reference).  It
 * pins the tiles it is made from between start() and stop(), and refuses to write
 * a slot that was given to another tile since the task was created.
 */
class CreateTile : public Task
{
public:
    TaskGraph *parent;          /* the graph holding this task's dependencies, or NULL */
    TileProducer *owner;
    int level, tx, ty;
    TileStorage::Slot *data;
    mutable void *cachedContext;
    bool initialized;           /* input tiles acquired (startCreateTile done, stopCreateTile not yet) */

    CreateTile(TileProducer *owner, int level, int tx, int ty, TileStorage::Slot *data, unsigned int deadline) :
        Task(owner->taskType, owner->isGpuProducer(), deadline), parent(NULL), owner(owner), level(level), tx(tx),
        ty(ty), data(data), cachedContext(NULL), initialized(true)
    {
        claim();
    }

    virtual ~CreateTile()
    {
        stop();
        if (owner != NULL) {
            owner->removeCreateTile(this);
            if (owner->cache != NULL) {
                owner->cache->createTileTaskDeleted(owner->getId(), level, tx, ty);
            }
        }
    }

    /* this task is the one that fills `data` */
    void claim()
    {
        data->lock(true);
        data->producerTask = this;
        data->lock(false);
    }

    virtual void *getContext() const
    {
        if (owner != NULL) {
            /* producer type + producer context: tasks that can share a batch */
            cachedContext = (void *) ((size_t) typeid(*owner).name() + (size_t) owner->getContext());
        }
        return cachedContext;
    }

    virtual void init(std::set<Task *> &initializedTasks)
    {
        (void) initializedTasks;
        if (!isDone() && owner != NULL) {
            start();
        }
    }

    void start()
    {
        if (initialized) {
            return;
        }
        if (parent != NULL) {
            parent->clearDependencies();
        }
        owner->startCreateTile(level, tx, ty, getDeadline(), this, parent);   /* may throw: nothing acquired then */
        if (parent != NULL) {
            /* tasks nobody needs any more (inputs of a previous incarnation) leave the graph */
            TaskGraph::TaskIterator i = parent->getLastTasks();
            while (i.hasNext()) {
                ptr<Task> t = i.next();
                if (t.get() != this) {
                    parent->removeTask(t);
                }
            }
        }
        initialized = true;
    }

    virtual void begin()
    {
        assert(!isDone());
        owner->beginCreateTile();
    }

    virtual bool run()
    {
        bool changes = true;
        assert(!isDone());
        data->lock(true);
        if (data->producerTask == this) {
            changes = owner->doCreateTile(level, tx, ty, data);
            data->id = TileCache::Tile::getTId(owner->getId(), level, tx, ty);
        }
        data->lock(false);
        return changes;
    }

    virtual void end()
    {
        owner->endCreateTile();
    }

    void stop()
    {
        if (initialized) {
            if (owner != NULL) {
                owner->stopCreateTile(level, tx, ty);
            }
            initialized = false;
        }
    }

    virtual void setIsDone(bool d, unsigned int t, reason r)
    {
        Task::setIsDone(d, t, r);
        if (d) {
            stop();
        } else if (r == DATA_NEEDED && owner != NULL) {
            /* called after the tile went back into the cache: it may sit in another slot now */
            TileCache::Tile *tile = owner->findTile(level, tx, ty, true);
            assert(tile != NULL);
            data = tile->data;
            claim();
            start();
        }
    }

    virtual const std::type_info *getTypeInfo()
    {
        return owner != NULL ? &typeid(*owner) : &typeid(*this);
    }
};

/* The graph of one CreateTile task and the tasks of its input tiles (TileProducer.cpp:255-316). */
class CreateTileTaskGraph : public TaskGraph
{
public:
    TileProducer *owner;
    CreateTile *root;

    CreateTileTaskGraph(TileProducer *owner) : TaskGraph(), owner(owner), root(NULL)
    {
    }

    virtual ~CreateTileTaskGraph()
    {
        if (root != NULL) {
            root->parent = NULL;
        }
        if (owner != NULL) {
            owner->removeCreateTile(this);
        }
    }

    /* a graph reused from TileCache's deleted-task map already holds its root and dependencies */
    void restore()
    {
        addTask(root);
    }
};

TileProducer::TileProducer(const char *type, const char *taskType, ptr<TileCache> cache, bool gpuProducer) :
    Object(type), taskType(taskType), gpuProducer(gpuProducer), id(-1), rootQuadSize(0.0f)
{
    init(cache, gpuProducer);
}

TileProducer::TileProducer(const char *type, const char *taskType) :
    Object(type), taskType(taskType), gpuProducer(false), id(-1), rootQuadSize(0.0f)
{
}

void TileProducer::init(ptr<TileCache> cache, bool gpuProducer)
{
    assert(cache != NULL);
    this->cache = cache;
    this->gpuProducer = gpuProducer;
    this->rootQuadSize = 0.0f;
    this->id = cache->nextProducerId++;
    cache->producers.insert(std::make_pair(id, this));
}

TileProducer::~TileProducer()
{
    assert(cache != NULL);
    cache->producers.erase(id);
    for (std::unordered_set<Task *>::iterator i = tasks.begin(); i != tasks.end(); ++i) {
        CreateTile *t = dynamic_cast<CreateTile *>(*i);
        if (t != NULL) {
            t->owner = NULL;
        } else {
            dynamic_cast<CreateTileTaskGraph *>(*i)->owner = NULL;
        }
    }
    layers.clear();
}

float TileProducer::getRootQuadSize()
{
    return rootQuadSize;
}

void TileProducer::setRootQuadSize(float size)
{
    rootQuadSize = size;
    for (size_t i = 0; i < layers.size(); ++i) {
        layers[i]->setCache(cache, id);
        layers[i]->setTileSize(cache->getStorage()->getTileSize(), getBorder(), getRootQuadSize());
    }
}

int TileProducer::getId()
{
    return id;
}

ptr<TileCache> TileProducer::getCache()
{
    return cache;
}

bool TileProducer::isGpuProducer()
{
    return gpuProducer;
}

int TileProducer::getBorder()
{
    return 0;
}

bool TileProducer::hasTile(int level, int tx, int ty)
{
    (void) level; (void) tx; (void) ty;
    return true;
}

bool TileProducer::hasChildren(int level, int tx, int ty)
{
    return hasTile(level + 1, 2 * tx, 2 * ty);
}

TileCache::Tile *TileProducer::findTile(int level, int tx, int ty, bool includeCache, bool done)
{
    TileCache::Tile *t = cache->findTile(id, level, tx, ty, includeCache);
    if (done && t != NULL && !t->task->isDone()) {
        t = NULL;
    }
    return t;
}

TileCache::Tile *TileProducer::getTile(int level, int tx, int ty, unsigned int deadline)
{
    int users = 0;
    TileCache::Tile *t = cache->getTile(id, level, tx, ty, deadline, &users);
    if (users == 0) {
        for (size_t i = 0; i < layers.size(); ++i) {
            layers[i]->useTile(level, tx, ty, deadline);
        }
    }
    return t;
}

bool TileProducer::prefetchTile(int level, int tx, int ty)
{
    if (cache->getScheduler() != NULL && cache->getScheduler()->supportsPrefetch(isGpuProducer())) {
        ptr<Task> task = cache->prefetchTile(id, level, tx, ty);
        if (task != NULL) {
            cache->getScheduler()->schedule(task);
            return true;
        }
    }
    for (size_t i = 0; i < layers.size(); ++i) {
        layers[i]->prefetchTile(level, tx, ty);
    }
    return false;
}

void TileProducer::putTile(TileCache::Tile *t)
{
    if (cache->putTile(t) == 0) {
        for (size_t i = 0; i < layers.size(); ++i) {
            layers[i]->unuseTile(t->level, t->tx, t->ty);
        }
    }
}

void TileProducer::invalidateTile(int level, int tx, int ty)
{
    getCache()->invalidateTile(getId(), level, tx, ty);
}

void TileProducer::invalidateTiles()
{
    cache->invalidateTiles(id);
}

void *TileProducer::getContext() const
{
    return NULL;
}

void TileProducer::getReferencedProducers(std::vector<ptr<TileProducer> > &producers) const
{
    (void) producers;
}

int TileProducer::getLayerCount() const
{
    return (int) layers.size();
}

ptr<TileLayer> TileProducer::getLayer(int index) const
{
    return layers[index];
}

bool TileProducer::hasLayers() const
{
    return layers.size() > 0;
}

void TileProducer::addLayer(ptr<TileLayer> l)
{
    layers.push_back(l);
}

ptr<Task> TileProducer::startCreateTile(int level, int tx, int ty, unsigned int deadline, ptr<Task> task,
                                        ptr<TaskGraph> owner)
{
    for (size_t i = 0; i < layers.size(); ++i) {
        layers[i]->startCreateTile(level, tx, ty, deadline, task, owner);
    }
    return owner == NULL ? task : owner.cast<Task>();
}

void TileProducer::beginCreateTile()
{
    for (size_t i = 0; i < layers.size(); ++i) {
        layers[i]->beginCreateTile();
    }
}

bool TileProducer::doCreateTile(int level, int tx, int ty, TileStorage::Slot *data)
{
    bool changes = false;
    for (size_t i = 0; i < layers.size(); ++i) {
        if (layers[i]->isEnabled()) {
            changes |= layers[i]->doCreateTile(level, tx, ty, data);
        }
    }
    return changes;
}

void TileProducer::endCreateTile()
{
    for (size_t i = 0; i < layers.size(); ++i) {
        layers[i]->endCreateTile();
    }
}

void TileProducer::stopCreateTile(int level, int tx, int ty)
{
    for (size_t i = 0; i < layers.size(); ++i) {
        layers[i]->stopCreateTile(level, tx, ty);
    }
}

ptr<Task> TileProducer::createTile(int level, int tx, int ty, TileStorage::Slot *data, unsigned int deadline,
                                   ptr<Task> old)
{
    assert(data != NULL);
    if (old != NULL) {
        ptr<CreateTileTaskGraph> r = old.cast<CreateTileTaskGraph>();
        if (r != NULL) {
            r->restore();
        } else {
            assert(old.cast<CreateTile>() != NULL);
        }
        return old;
    }
    ptr<CreateTile> t = new CreateTile(this, level, tx, ty, data, deadline);
    ptr<Task> r;
    try {
        r = startCreateTile(level, tx, ty, deadline, t, NULL);
    } catch (...) {
        t->initialized = false;   /* startCreateTile released what it had acquired */
        throw;
    }
    std::lock_guard<std::mutex> lock(mutex);
    tasks.insert(t.get());
    if (r.get() != t.get()) {
        assert(r.cast<CreateTileTaskGraph>() != NULL);
        tasks.insert(r.get());
    }
    return r;
}

ptr<TaskGraph> TileProducer::createTaskGraph(ptr<Task> task)
{
    ptr<CreateTile> t = task.cast<CreateTile>();
    assert(t != NULL);
    ptr<CreateTileTaskGraph> r = new CreateTileTaskGraph(this);
    r->addTask(t);
    r->root = t.get();
    t->parent = r.get();
    return r;
}

void TileProducer::cacheFull(const char *producerType)
{
    const std::string msg = std::string("Insufficient tile cache size (") + producerType + ")";
    if (Logger::ERROR_LOGGER != NULL) {
        Logger::ERROR_LOGGER->log("CACHE", msg);
    }
    throw CacheFullError(msg);
}

void TileProducer::removeCreateTile(Task *t)
{
    std::lock_guard<std::mutex> lock(mutex);
    tasks.erase(t);
}

}  // namespace proland
