#include "proland/dem/ResidualProducer.h"

#include <algorithm>
#include <cassert>
#include <cstring>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace proland
{

ResidualProducer::ResidualProducer(ptr<TileCache> cache, const char *name, int deltaLevel, float zscale) :
    TileProducer("ResidualProducer", "CreateResidualTile"), fileData(NULL), fileSize(0), storage(NULL), tileCount(0)
{
    init(cache, name, deltaLevel, zscale);
}

ResidualProducer::ResidualProducer() :
    TileProducer("ResidualProducer", "CreateResidualTile"), fileData(NULL), fileSize(0), storage(NULL), tileCount(0)
{
}

void ResidualProducer::init(ptr<TileCache> cache, const char *name, int deltaLevel, float zscale)
{
    /* a CPU producer in the reference (worker threads); the decode is device work here */
    TileProducer::init(cache, true);
    this->name = name;
    storage = dynamic_cast<GPUTileStorage *>(cache->getStorage().get());
    if (storage == NULL || storage->getInternalFormat() != R32F) {
        if (Logger::ERROR_LOGGER != NULL) {
            Logger::ERROR_LOGGER->log("DEM", "ResidualProducer needs a cpuFloatTileStorage (device float pool)");
        }
        throw std::invalid_argument("ResidualProducer: bad tile storage");
    }
    context = storage->getContext();
    BatchSourceRegistration registration(context.get(), this);

    if (strlen(name) == 0) {
        /* no file: all-zero residuals of any level (ResidualProducer.cpp:76-84) */
        this->tileSize = storage->getTileSize() - 5;
        this->minLevel = 0;
        this->maxLevel = 32;
        this->rootLevel = 0;
        this->deltaLevel = 0;
        this->rootTx = 0;
        this->rootTy = 0;
        this->scale = 1.0;
        this->header = 0;
        registration.commit();
        return;
    }

    this->minLevel = 0;
    this->maxLevel = -1;
    this->tileSize = storage->getTileSize() - 5;
    this->rootLevel = 0;
    this->rootTx = 0;
    this->rootTy = 0;
    this->scale = 1.0;
    const int fd = open(name, O_RDONLY);
    struct stat st;
    if (fd < 0 || fstat(fd, &st) != 0 || st.st_size < 28) {
        if (fd >= 0) close(fd);
        if (Logger::ERROR_LOGGER != NULL) {
            Logger::ERROR_LOGGER->log("DEM", "Cannot open file '" + std::string(name) + "'");
        }
        /* like the reference: maxLevel = -1, the producer has no tile */
    } else {
        void *map = mmap(NULL, (size_t) st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
        close(fd);
        if (map == MAP_FAILED) {
            throw DeviceError(PL_ERR_IO, "cannot map '" + std::string(name) + "'");
        }
        fileData = static_cast<const unsigned char *>(map);
        fileSize = (size_t) st.st_size;
        int head[6];
        memcpy(head, fileData, sizeof(head));
        memcpy(&scale, fileData + sizeof(head), sizeof(float));
        minLevel = head[0];
        maxLevel = head[1];
        tileSize = head[2];
        rootLevel = head[3];
        rootTx = head[4];
        rootTy = head[5];
    }

    this->deltaLevel = rootLevel == 0 ? deltaLevel : 0;
    scale = scale * zscale;

    const int ntiles = minLevel + ((1 << (std::max(maxLevel - minLevel, 0) * 2 + 2)) - 1) / 3;
    header = sizeof(float) + sizeof(int) * (6 + ntiles * 2);
    offsets.assign((size_t) ntiles * 2, 0u);
    if (fileData != NULL) {
        if (fileSize < header) {
            throw DeviceError(PL_ERR_CORRUPT, "'" + std::string(name) + "': offset table is truncated");
        }
        memcpy(&offsets[0], fileData + 28, sizeof(unsigned int) * ntiles * 2);
        if (tileSize + 5 != storage->getTileSize()) {
            throw std::invalid_argument("ResidualProducer: the storage tile size must be the file's tile size + 5");
        }
    }
    assert(fileData == NULL || this->deltaLevel <= minLevel);
    registration.commit();
}

ResidualProducer::~ResidualProducer()
{
    if (context != NULL) {
        context->removeSource(this);
    }
    if (fileData != NULL) {
        munmap(const_cast<unsigned char *>(fileData), fileSize);
    }
}

int ResidualProducer::getBorder()
{
    return 2;
}

int ResidualProducer::getMinLevel()
{
    return minLevel;
}

int ResidualProducer::getDeltaLevel()
{
    return deltaLevel;
}

void ResidualProducer::addProducer(ptr<ResidualProducer> p)
{
    producers.push_back(p);
}

bool ResidualProducer::hasTile(int level, int tx, int ty)
{
    const int l = level + deltaLevel - rootLevel;
    if (l >= 0 && (tx >> l) == rootTx && (ty >> l) == rootTy) {
        if (l <= maxLevel) {
            return true;
        }
        for (size_t i = 0; i < producers.size(); ++i) {
            if (producers[i]->hasTile(level + deltaLevel, tx, ty)) {
                return true;
            }
        }
    }
    return false;
}

int ResidualProducer::getTileSize(int level)
{
    return level < minLevel ? tileSize >> (minLevel - level) : tileSize;
}

int ResidualProducer::getTileId(int level, int tx, int ty)
{
    if (level < minLevel) {
        return level;
    }
    const int l = std::max(level - minLevel, 0);
    return minLevel + tx + ty * (1 << l) + ((1 << (2 * l)) - 1) / 3;
}

bool ResidualProducer::doCreateTile(int level, int tx, int ty, TileStorage::Slot *data)
{
    const int l = level + deltaLevel - rootLevel;
    if (l >= 0 && (tx >> l) == rootTx && (ty >> l) == rootTy) {
        if (l > maxLevel) {
            for (size_t i = 0; i < producers.size(); ++i) {
                producers[i]->doCreateTile(level + deltaLevel, tx, ty, data);
            }
            return true;
        }
    } else {
        return true;
    }

    if (Logger::DEBUG_LOGGER != NULL) {
        Logger::DEBUG_LOGGER->logf("DEM", "Residual tile %d %d %d %d", getId(), level, tx, ty);
    }

    GPUTileStorage::GPUSlot *slot = dynamic_cast<GPUTileStorage::GPUSlot *>(data);
    assert(slot != NULL);
    if (data->getOwner() != storage) {
        /* a nested producer writes the slot its parent was given: they must share the storage */
        throw std::logic_error("ResidualProducer: nested residual producers must use the cache of their parent");
    }
    Job j;
    j.level = l;
    j.tx = tx - (rootTx << l);
    j.ty = ty - (rootTy << l);
    j.slot = slot->l;
    j.root = deltaLevel > 0 && l == deltaLevel;
    pending.push_back(j);
    ++tileCount;
    return true;
}

void ResidualProducer::endCreateTile()
{
    TileProducer::endCreateTile();
    if (!context->inBatch()) {
        context->flush();
    }
}

void ResidualProducer::blobOf(int tileid, uint64_t *offset, uint32_t *size) const
{
    if (tileid < 0 || (size_t) (2 * tileid + 1) >= offsets.size()) {
        throw DeviceError(PL_ERR_CORRUPT, "'" + name + "': tile id out of range");
    }
    const uint64_t a = (uint64_t) header + offsets[2 * tileid], b = (uint64_t) header + offsets[2 * tileid + 1];
    if (b < a || b > fileSize) {
        throw DeviceError(PL_ERR_CORRUPT, "'" + name + "': blob outside the file");
    }
    *offset = a;
    *size = (uint32_t) (b - a);
}

void ResidualProducer::decodeOne(int level, int tx, int ty, int outSlot, int addSlot)
{
    uint64_t off;
    uint32_t size;
    blobOf(getTileId(level, tx, ty), &off, &size);
    const int32_t w = getTileSize(level) + 5, os = outSlot, as = addSlot;
    DeviceContext::check(pl_residual_decode_batch(context->handle(), storage->getPool(), 1, fileData, &off, &size, &w, &os,
                                                  addSlot == -1 ? NULL : &as, scale));
}

void ResidualProducer::flushBatch()
{
    if (pending.empty()) {
        return;
    }
    std::vector<Job> jobs;
    jobs.swap(pending);
    pl_pool *pool = storage->getPool();

    if (name.empty()) {
        /* readTile without a file writes zeros (ResidualProducer.cpp:281-287) */
        const int w = storage->getTileSize();
        std::vector<float> zeros((size_t) w * w, 0.0f);
        for (size_t i = 0; i < jobs.size(); ++i) {
            DeviceContext::check(pl_pool_upload(pool, jobs[i].slot, &zeros[0], zeros.size() * sizeof(float)));
        }
        return;
    }

    /* plain tiles: one batched decode straight out of the mapped file */
    std::vector<uint64_t> offs;
    std::vector<uint32_t> sizes;
    std::vector<int32_t> widths, slots;
    for (size_t i = 0; i < jobs.size(); ++i) {
        if (jobs[i].root) continue;
        uint64_t off;
        uint32_t size;
        blobOf(getTileId(jobs[i].level, jobs[i].tx, jobs[i].ty), &off, &size);
        offs.push_back(off);
        sizes.push_back(size);
        widths.push_back(getTileSize(jobs[i].level) + 5);
        slots.push_back(jobs[i].slot);
    }
    if (!offs.empty()) {
        DeviceContext::check(pl_residual_decode_batch(context->handle(), pool, (int) offs.size(), fileData, &offs[0], &sizes[0],
                                                      &widths[0], &slots[0], NULL, scale));
    }
    /* root tiles: stored level 0, then `delta` times upsample + add the stored residual of that level */
    for (size_t i = 0; i < jobs.size(); ++i) {
        if (!jobs[i].root) continue;
        decodeOne(0, 0, 0, jobs[i].slot, -1);
        for (int k = 1; k <= deltaLevel; ++k) {
            DeviceContext::check(pl_residual_upsample(context->handle(), pool, jobs[i].slot, PL_SLOT_SCRATCH, getTileSize(k), 0, 0));
            decodeOne(k, 0, 0, jobs[i].slot, PL_SLOT_SCRATCH);
        }
    }
}

}  // namespace proland
