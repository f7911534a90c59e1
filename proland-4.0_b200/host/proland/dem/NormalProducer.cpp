#include "proland/dem/NormalProducer.h"

#include <cassert>

namespace proland
{

NormalProducer::NormalProducer(ptr<TileCache> cache, ptr<TileProducer> elevationTiles, int gridSize, bool deform) :
    TileProducer("NormalProducer", "CreateNormalTile"), storage(NULL), elevationStorage(NULL), tileCount(0), batchCount(0)
{
    init(cache, elevationTiles, gridSize, deform);
}

NormalProducer::NormalProducer() :
    TileProducer("NormalProducer", "CreateNormalTile"), deform(false), gridMeshSize(24), storage(NULL),
    elevationStorage(NULL), tileCount(0), batchCount(0)
{
}

void NormalProducer::init(ptr<TileCache> cache, ptr<TileProducer> elevationTiles, int gridSize, bool deform)
{
    TileProducer::init(cache, true);
    this->elevationTiles = elevationTiles;
    this->deform = deform;
    this->gridMeshSize = gridSize;
    storage = dynamic_cast<GPUTileStorage *>(cache->getStorage().get());
    elevationStorage = dynamic_cast<GPUTileStorage *>(elevationTiles->getCache()->getStorage().get());
    if (storage == NULL || elevationStorage == NULL ||
        (storage->getInternalFormat() != RG8 && storage->getInternalFormat() != RGBA8)) {
        if (Logger::ERROR_LOGGER != NULL) {
            Logger::ERROR_LOGGER->log("DEM", "NormalProducer needs an RG8 or RGBA8 gpuTileStorage and GPU elevation tiles");
        }
        throw std::invalid_argument("NormalProducer: bad tile storage");
    }
    /* the reference asserts these (NormalProducer.cpp:101-104) */
    if (storage->getTileSize() != elevationStorage->getTileSize() - 2 * elevationTiles->getBorder() ||
        (storage->getTileSize() - 1) % gridSize != 0 || storage->getContext() != elevationStorage->getContext()) {
        throw std::invalid_argument("NormalProducer: normal tile size must be the elevation tile size minus its borders, "
                                    "(tileSize - 1) a multiple of gridSize, same device");
    }
    context = storage->getContext();
    context->addSource(this);
}

NormalProducer::~NormalProducer()
{
    if (context != NULL) {
        context->removeSource(this);
    }
}

void NormalProducer::getReferencedProducers(std::vector<ptr<TileProducer> > &producers) const
{
    producers.push_back(elevationTiles);
}

void NormalProducer::setRootQuadSize(float size)
{
    TileProducer::setRootQuadSize(size);
    elevationTiles->setRootQuadSize(size);
}

int NormalProducer::getBorder()
{
    return 0;
}

bool NormalProducer::hasTile(int level, int tx, int ty)
{
    return elevationTiles->hasTile(level, tx, ty);
}

void *NormalProducer::getContext() const
{
    return storage;
}

ptr<Task> NormalProducer::startCreateTile(int level, int tx, int ty, unsigned int deadline, ptr<Task> task,
                                          ptr<TaskGraph> owner)
{
    ptr<TaskGraph> result = owner == NULL ? createTaskGraph(task) : owner;
    TileCache::Tile *parentTile = NULL;
    if (level > 0) {
        parentTile = getTile(level - 1, tx / 2, ty / 2, deadline);
        if (parentTile == NULL) {
            cacheFull("NormalProducer");
        }
        result->addTask(parentTile->task);
        result->addDependency(task, parentTile->task);
    }
    TileCache::Tile *t = NULL;
    try {
        t = elevationTiles->getTile(level, tx, ty, deadline);
    } catch (...) {
        if (parentTile != NULL) putTile(parentTile);
        throw;
    }
    if (t == NULL) {
        if (parentTile != NULL) putTile(parentTile);
        cacheFull("ElevationProducer");
    }
    result->addTask(t->task);
    result->addDependency(task, t->task);
    return result;
}

void NormalProducer::beginCreateTile()
{
}

pl_norm_scene NormalProducer::scene() const
{
    pl_norm_scene sc = pl_norm_scene();
    sc.tile_w = storage->getTileSize();
    sc.grid = (storage->getTileSize() - 1) / gridMeshSize;
    sc.elev_border = elevationTiles->getBorder();
    sc.elev_filter = elevationStorage->getFilter() == LINEAR ? PL_FILTER_LINEAR : PL_FILTER_NEAREST;
    sc.parent_filter = storage->getFilter() == LINEAR ? PL_FILTER_LINEAR : PL_FILTER_NEAREST;
    sc.sphere = deform ? 1 : 0;
    return sc;
}

bool NormalProducer::doCreateTile(int level, int tx, int ty, TileStorage::Slot *data)
{
    if (Logger::DEBUG_LOGGER != NULL) {
        Logger::DEBUG_LOGGER->logf("DEM", "Normal tile %d %d %d %d", getId(), level, tx, ty);
    }
    GPUTileStorage::GPUSlot *gpuData = dynamic_cast<GPUTileStorage::GPUSlot *>(data);
    assert(gpuData != NULL);
    const int components = storage->getComponents();

    const pl_norm_scene sc = scene();
    pl_norm_req req;
    pl_norm_make_req(&sc, (double) getRootQuadSize(), components, level, tx, ty, &req);
    req.out_slot = gpuData->l;

    if (level > 0 && components == 4) {
        TileCache::Tile *t = findTile(level - 1, tx / 2, ty / 2);
        assert(t != NULL);
        GPUTileStorage::GPUSlot *parentGpuData = dynamic_cast<GPUTileStorage::GPUSlot *>(t->getData());
        assert(parentGpuData != NULL);
        req.parent_slot = parentGpuData->l;
    }
    TileCache::Tile *t = elevationTiles->findTile(level, tx, ty);
    assert(t != NULL);
    GPUTileStorage::GPUSlot *elevationGpuData = dynamic_cast<GPUTileStorage::GPUSlot *>(t->getData());
    assert(elevationGpuData != NULL);
    req.elev_slot = elevationGpuData->l;

    pending.push_back(req);
    ++tileCount;
    return true;
}

void NormalProducer::endCreateTile()
{
    if (!context->inBatch()) {
        context->flush();
    }
}

void NormalProducer::flushBatch()
{
    if (pending.empty()) {
        return;
    }
    const pl_norm_scene sc = scene();
    std::vector<pl_norm_req> batch;
    batch.swap(pending);
    ++batchCount;
    DeviceContext::check(pl_normal_batch(context->handle(), &sc, storage->getPool(), elevationStorage->getPool(),
                                         (int) batch.size(), &batch[0]));
}

void NormalProducer::stopCreateTile(int level, int tx, int ty)
{
    if (level > 0) {
        TileCache::Tile *t = findTile(level - 1, tx / 2, ty / 2);
        assert(t != NULL);
        putTile(t);
    }
    TileCache::Tile *t = elevationTiles->findTile(level, tx, ty);
    assert(t != NULL);
    elevationTiles->putTile(t);
}

}  // namespace proland
