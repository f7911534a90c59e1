#include "proland/ortho/OrthoProducer.h"

#include <cassert>
#include <cstring>

namespace proland
{

OrthoProducer::OrthoProducer(ptr<TileCache> cache, ptr<TileProducer> residualTiles, const float rootNoiseColor[4],
                             const float noiseColor[4], std::vector<float> &noiseAmp, bool noiseHsv, float scale,
                             int maxLevel, int face) :
    TileProducer("OrthoProducer", "CreateOrthoTile"), storage(NULL), tileCount(0), batchCount(0)
{
    init(cache, residualTiles, rootNoiseColor, noiseColor, noiseAmp, noiseHsv, scale, maxLevel, face);
}

OrthoProducer::OrthoProducer() :
    TileProducer("OrthoProducer", "CreateOrthoTile"), face(1), maxLevel(-1), storage(NULL), tileCount(0), batchCount(0)
{
    memset(&scene, 0, sizeof(scene));
}

void OrthoProducer::init(ptr<TileCache> cache, ptr<TileProducer> residualTiles, const float rootNoiseColor[4],
                         const float noiseColor[4], std::vector<float> &noiseAmp, bool noiseHsv, float scale,
                         int maxLevel, int face)
{
    TileProducer::init(cache, true);
    storage = dynamic_cast<GPUTileStorage *>(cache->getStorage().get());
    if (storage == NULL || (storage->getInternalFormat() != RGBA8 && storage->getInternalFormat() != RGB8)) {
        if (Logger::ERROR_LOGGER != NULL) {
            Logger::ERROR_LOGGER->log("ORTHO", "OrthoProducer needs an RGB8 / RGBA8 gpuTileStorage");
        }
        throw std::invalid_argument("OrthoProducer: bad tile storage");
    }
    const int tileWidth = storage->getTileSize();
    if ((tileWidth - 4) % 8 != 0) {
        throw std::invalid_argument("OrthoProducer: (tileSize - 4) must be a multiple of 8");
    }
    if (noiseAmp.size() > 32) {
        throw std::invalid_argument("OrthoProducer: at most 32 noise amplitudes");
    }
    this->residualTiles = residualTiles;
    this->face = face;
    this->maxLevel = maxLevel;
    this->context = storage->getContext();
    memset(&scene, 0, sizeof(scene));
    scene.tile_w = tileWidth;
    scene.channels = storage->getComponents();
    scene.hsv = noiseHsv ? 1 : 0;
    scene.face = face;
    scene.scale = scale;
    for (int i = 0; i < 4; ++i) {
        scene.noise_color[i] = noiseColor[i];
        scene.root_noise_color[i] = rootNoiseColor[i];
    }
    scene.n_amp = (int) noiseAmp.size();
    scene.max_level = maxLevel;
    scene.out_channels = storage->getComponents();
    for (size_t i = 0; i < noiseAmp.size(); ++i) {
        scene.noise_amp[i] = noiseAmp[i];
    }
    /* orthoNoiseFactory->get(tileWidth), OrthoProducer.cpp:155 */
    context->ensureOrthoNoise(tileWidth);
    BatchSourceRegistration registration(context.get(), this);
    if (residualTiles != NULL) {
        GPUTileStorage *rs = dynamic_cast<GPUTileStorage *>(residualTiles->getCache()->getStorage().get());
        if (rs == NULL || (rs->getInternalFormat() != RGBA8 && rs->getInternalFormat() != RGB8) ||
            rs->getContext() != context || rs->getTileSize() != tileWidth) {
            throw std::invalid_argument("OrthoProducer: residual tiles must live in a byte storage of the same device and tile size");
        }
        /* channels = the residual storage's (OrthoProducer.cpp:168-172) */
        scene.channels = rs->getComponents();
        assert(storage->getComponents() >= scene.channels);
    }
    registration.commit();
}

OrthoProducer::~OrthoProducer()
{
    if (context != NULL) {
        context->removeSource(this);
    }
}

void OrthoProducer::getReferencedProducers(std::vector<ptr<TileProducer> > &producers) const
{
    if (residualTiles != NULL) {
        producers.push_back(residualTiles);
    }
}

void OrthoProducer::setRootQuadSize(float size)
{
    TileProducer::setRootQuadSize(size);
    if (residualTiles != NULL) {
        residualTiles->setRootQuadSize(size);
    }
}

int OrthoProducer::getBorder()
{
    assert(residualTiles == NULL || residualTiles->getBorder() == 2);
    return 2;
}

bool OrthoProducer::hasTile(int level, int tx, int ty)
{
    (void) tx;
    (void) ty;
    return maxLevel == -1 || level <= maxLevel;
}

void *OrthoProducer::getContext() const
{
    return storage;
}

bool OrthoProducer::prefetchTile(int level, int tx, int ty)
{
    bool b = TileProducer::prefetchTile(level, tx, ty);
    if (!b) {
        if (residualTiles != NULL && residualTiles->hasTile(level, tx, ty)) {
            residualTiles->prefetchTile(level, tx, ty);
        }
    }
    return b;
}

ptr<Task> OrthoProducer::startCreateTile(int level, int tx, int ty, unsigned int deadline, ptr<Task> task,
                                         ptr<TaskGraph> owner)
{
    ptr<TaskGraph> result = owner == NULL ? createTaskGraph(task) : owner;
    TileCache::Tile *parentTile = NULL;
    if (level > 0) {
        parentTile = getTile(level - 1, tx / 2, ty / 2, deadline);
        if (parentTile == NULL) {
            cacheFull("OrthoProducer");
        }
        result->addTask(parentTile->task);
        result->addDependency(task, parentTile->task);
    }
    if (residualTiles != NULL && residualTiles->hasTile(level, tx, ty)) {
        TileCache::Tile *t = residualTiles->getTile(level, tx, ty, deadline);
        if (t == NULL) {
            if (parentTile != NULL) putTile(parentTile);
            cacheFull("OrthoProducer residuals");
        }
        result->addTask(t->task);
        result->addDependency(task, t->task);
    }
    TileProducer::startCreateTile(level, tx, ty, deadline, task, result);
    return result;
}

void OrthoProducer::beginCreateTile()
{
    TileProducer::beginCreateTile();
}

bool OrthoProducer::doCreateTile(int level, int tx, int ty, TileStorage::Slot *data)
{
    if (Logger::DEBUG_LOGGER != NULL) {
        Logger::DEBUG_LOGGER->logf("ORTHO", "Ortho tile %d %d %d %d", getId(), level, tx, ty);
    }
    GPUTileStorage::GPUSlot *gpuData = dynamic_cast<GPUTileStorage::GPUSlot *>(data);
    assert(gpuData != NULL);
    if (hasLayers()) {
        if (Logger::ERROR_LOGGER != NULL) {
            Logger::ERROR_LOGGER->log("ORTHO", "ortho layers are not part of the device path");
        }
        throw std::logic_error("OrthoProducer: layers are not supported");
    }
    const bool hasResidual = residualTiles != NULL && residualTiles->hasTile(level, tx, ty);
    /* tileWidth, coarseLevelOSL, residualOSH, noiseUVLH, noiseColor (OrthoProducer.cpp:286-366) */
    pl_ortho_req req;
    pl_ortho_make_req(&scene, level, tx, ty, hasResidual ? 1 : 0, &req);
    req.out_slot = gpuData->l;
    if (level > 0) {
        TileCache::Tile *t = findTile(level - 1, tx / 2, ty / 2);
        assert(t != NULL);
        GPUTileStorage::GPUSlot *parentGpuData = dynamic_cast<GPUTileStorage::GPUSlot *>(t->getData());
        assert(parentGpuData != NULL);
        req.parent_slot = parentGpuData->l;
    }
    if (hasResidual) {
        TileCache::Tile *t = residualTiles->findTile(level, tx, ty);
        assert(t != NULL);
        GPUTileStorage::GPUSlot *residual = dynamic_cast<GPUTileStorage::GPUSlot *>(t->getData());
        assert(residual != NULL);
        req.resid_slot = residual->l;
    }
    pending.push_back(req);
    ++tileCount;
    return true;
}

void OrthoProducer::endCreateTile()
{
    TileProducer::endCreateTile();
    if (!context->inBatch()) {
        context->flush();
    }
}

void OrthoProducer::flushBatch()
{
    if (pending.empty()) {
        return;
    }
    pl_pool *resid = NULL;
    if (residualTiles != NULL) {
        resid = static_cast<GPUTileStorage *>(residualTiles->getCache()->getStorage().get())->getPool();
    }
    std::vector<pl_ortho_req> batch;
    batch.swap(pending);
    ++batchCount;
    DeviceContext::check(pl_ortho_batch(context->handle(), &scene, storage->getPool(), resid, (int) batch.size(), &batch[0]));
}

void OrthoProducer::stopCreateTile(int level, int tx, int ty)
{
    if (level > 0) {
        TileCache::Tile *t = findTile(level - 1, tx / 2, ty / 2);
        assert(t != NULL);
        putTile(t);
    }
    if (residualTiles != NULL && residualTiles->hasTile(level, tx, ty)) {
        TileCache::Tile *t = residualTiles->findTile(level, tx, ty);
        assert(t != NULL);
        residualTiles->putTile(t);
    }
    TileProducer::stopCreateTile(level, tx, ty);
}

}  // namespace proland
