#include "proland/dem/ElevationProducer.h"

#include <cassert>

namespace proland
{

ElevationProducer::ElevationProducer(ptr<TileCache> cache, ptr<TileProducer> residualTiles, int gridMeshSize,
                                     std::vector<float> &noiseAmp, bool flipDiagonals, int face, Variant variant) :
    TileProducer("ElevationProducer", "CreateElevationTile"), storage(NULL), tileCount(0), batchCount(0)
{
    init(cache, residualTiles, gridMeshSize, noiseAmp, flipDiagonals, face, variant);
}

ElevationProducer::ElevationProducer() :
    TileProducer("ElevationProducer", "CreateElevationTile"), face(0), gridMeshSize(24), flipDiagonals(false),
    storage(NULL), tileCount(0), batchCount(0)
{
}

void ElevationProducer::init(ptr<TileCache> cache, ptr<TileProducer> residualTiles, int gridMeshSize,
                             std::vector<float> &noiseAmp, bool flipDiagonals, int face, Variant variant)
{
    TileProducer::init(cache, true);
    storage = dynamic_cast<GPUTileStorage *>(cache->getStorage().get());
    if (storage == NULL || storage->getComponents() < 3) {
        if (Logger::ERROR_LOGGER != NULL) {
            Logger::ERROR_LOGGER->log("DEM", "ElevationProducer needs an RGB32F / RGBA32F gpuTileStorage");
        }
        throw std::invalid_argument("ElevationProducer: bad tile storage");
    }
    const int tileWidth = storage->getTileSize();
    if ((tileWidth - 5) % gridMeshSize != 0) {
        throw std::invalid_argument("ElevationProducer: (tileSize - 5) must be a multiple of gridSize");
    }
    this->residualTiles = residualTiles;
    this->noiseAmp = noiseAmp;
    this->gridMeshSize = gridMeshSize;
    this->flipDiagonals = flipDiagonals;
    this->face = face;
    this->variant = variant;
    this->context = storage->getContext();
    /* demNoiseFactory->get(tileWidth), ElevationProducer.cpp:176 */
    context->ensureNoise(tileWidth);
    BatchSourceRegistration registration(context.get(), this);
    if (residualTiles != NULL) {
        GPUTileStorage *rs = dynamic_cast<GPUTileStorage *>(residualTiles->getCache()->getStorage().get());
        if (rs == NULL || rs->getInternalFormat() != R32F || rs->getContext() != context) {
            throw std::invalid_argument("ElevationProducer: residual tiles must live in a float storage of the same device");
        }
    }
    registration.commit();
}

ElevationProducer::~ElevationProducer()
{
    if (context != NULL) {
        context->removeSource(this);
    }
}

void ElevationProducer::getReferencedProducers(std::vector<ptr<TileProducer> > &producers) const
{
    if (residualTiles != NULL) {
        producers.push_back(residualTiles);
    }
}

void ElevationProducer::setRootQuadSize(float size)
{
    TileProducer::setRootQuadSize(size);
    if (residualTiles != NULL) {
        residualTiles->setRootQuadSize(size);
    }
}

int ElevationProducer::getBorder()
{
    assert(residualTiles == NULL || residualTiles->getBorder() == 2);
    return 2;
}

void *ElevationProducer::getContext() const
{
    return storage;
}

int ElevationProducer::residualMod() const
{
    const int tileSize = storage->getTileSize() - 5;
    const int residualTileSize = residualTiles->getCache()->getStorage()->getTileSize() - 5;
    return residualTileSize / tileSize;
}

ptr<Task> ElevationProducer::startCreateTile(int level, int tx, int ty, unsigned int deadline, ptr<Task> task,
                                             ptr<TaskGraph> owner)
{
    ptr<TaskGraph> result = owner == NULL ? createTaskGraph(task) : owner;
    TileCache::Tile *parentTile = NULL;
    if (level > 0) {
        parentTile = getTile(level - 1, tx / 2, ty / 2, deadline);
        if (parentTile == NULL) {
            cacheFull("ElevationProducer");
        }
        result->addTask(parentTile->task);
        result->addDependency(task, parentTile->task);
    }
    if (residualTiles != NULL) {
        const int mod = residualMod();
        if (residualTiles->hasTile(level, tx / mod, ty / mod)) {
            TileCache::Tile *t = residualTiles->getTile(level, tx / mod, ty / mod, deadline);
            if (t == NULL) {
                if (parentTile != NULL) putTile(parentTile);
                cacheFull("ResidualProducer");
            }
            result->addTask(t->task);
            result->addDependency(task, t->task);
        }
    }
    TileProducer::startCreateTile(level, tx, ty, deadline, task, result);
    return result;
}

void ElevationProducer::beginCreateTile()
{
    TileProducer::beginCreateTile();
}

bool ElevationProducer::doCreateTile(int level, int tx, int ty, TileStorage::Slot *data)
{
    if (Logger::DEBUG_LOGGER != NULL) {
        Logger::DEBUG_LOGGER->logf("DEM", "Elevation tile %d %d %d %d", getId(), level, tx, ty);
    }
    GPUTileStorage::GPUSlot *gpuData = dynamic_cast<GPUTileStorage::GPUSlot *>(data);
    assert(gpuData != NULL);
    if (hasLayers()) {
        if (Logger::ERROR_LOGGER != NULL) {
            Logger::ERROR_LOGGER->log("DEM", "elevation layers (blendShader pass) are not part of the device path");
        }
        throw std::logic_error("ElevationProducer: layers are not supported");
    }
    const int tileWidth = data->getOwner()->getTileSize();

    int residualTileWidth = 0;
    bool hasResidual = false;
    int mod = 1;
    if (residualTiles != NULL) {
        residualTileWidth = residualTiles->getCache()->getStorage()->getTileSize();
        mod = residualMod();
        hasResidual = residualTiles->hasTile(level, tx / mod, ty / mod);
    }

    /* tileWSDF, coarseLevelOSL, residualOSH, noiseUVLH (ElevationProducer.cpp:305-376) */
    pl_elev_req req;
    pl_elev_make_req(tileWidth, getRootQuadSize(), noiseAmp.empty() ? NULL : &noiseAmp[0], (int) noiseAmp.size(), face,
                     level, tx, ty, residualTileWidth, hasResidual ? 1 : 0, &req);
    req.out_slot = gpuData->l;
    if (level > 0) {
        TileCache::Tile *t = findTile(level - 1, tx / 2, ty / 2);
        assert(t != NULL);
        GPUTileStorage::GPUSlot *parentGpuData = dynamic_cast<GPUTileStorage::GPUSlot *>(t->getData());
        assert(parentGpuData != NULL);
        req.parent_slot = parentGpuData->l;
    }
    if (hasResidual) {
        TileCache::Tile *t = residualTiles->findTile(level, tx / mod, ty / mod);
        assert(t != NULL);
        GPUTileStorage::GPUSlot *residual = dynamic_cast<GPUTileStorage::GPUSlot *>(t->getData());
        assert(residual != NULL);
        req.resid_slot = residual->l;
    }
    pending.push_back(req);
    ++tileCount;
    return true;
}

void ElevationProducer::endCreateTile()
{
    TileProducer::endCreateTile();
    if (!context->inBatch()) {
        /* a scheduler that runs one task at a time expects the tile to exist now */
        context->flush();
    }
}

void ElevationProducer::flushBatch()
{
    if (pending.empty()) {
        return;
    }
    pl_elev_scene scene;
    scene.tile_w = storage->getTileSize();
    scene.grid = (storage->getTileSize() - 5) / gridMeshSize;
    scene.flip = flipDiagonals ? 1 : 0;
    scene.noise_mode = variant.slopeNoise ? PL_NOISE_SLOPE : PL_NOISE_PLAIN;
    scene.no_clamp = variant.noClamp ? 1 : 0;
    scene.want_stats = 1;
    scene.resid_scale = 1.0f;
    scene.pad_ = 0;
    pl_pool *resid = NULL;
    if (residualTiles != NULL) {
        resid = static_cast<GPUTileStorage *>(residualTiles->getCache()->getStorage().get())->getPool();
    }
    std::vector<pl_elev_req> batch;
    batch.swap(pending);
    ++batchCount;
    DeviceContext::check(pl_elevation_batch(context->handle(), &scene, storage->getPool(), resid, (int) batch.size(), &batch[0]));
}

void ElevationProducer::stopCreateTile(int level, int tx, int ty)
{
    if (level > 0) {
        TileCache::Tile *t = findTile(level - 1, tx / 2, ty / 2);
        assert(t != NULL);
        putTile(t);
    }
    if (residualTiles != NULL) {
        const int mod = residualMod();
        if (residualTiles->hasTile(level, tx / mod, ty / mod)) {
            TileCache::Tile *t = residualTiles->findTile(level, tx / mod, ty / mod);
            assert(t != NULL);
            residualTiles->putTile(t);
        }
    }
    TileProducer::stopCreateTile(level, tx, ty);
}

void ElevationProducer::getTileMinMax(TileCache::Tile *t, float *zmin, float *zmax)
{
    GPUTileStorage::GPUSlot *s = dynamic_cast<GPUTileStorage::GPUSlot *>(t->getData());
    assert(s != NULL);
    context->flush();
    float mm[2];
    const int32_t slot = s->l;
    DeviceContext::check(pl_elev_stats_download(context->handle(), storage->getPool(), 1, &slot, mm));
    *zmin = mm[0];
    *zmax = mm[1];
}

}  // namespace proland
