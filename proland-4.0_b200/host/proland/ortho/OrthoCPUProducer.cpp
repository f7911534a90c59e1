#include "proland/ortho/OrthoCPUProducer.h"

#include <cassert>
#include <cstring>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace proland
{

OrthoCPUProducer::OrthoCPUProducer(ptr<TileCache> cache, const char *name) :
    TileProducer("OrthoCPUProducer", "CreateOrthoCPUTile"), fileData(NULL), fileSize(0), storage(NULL), tileCount(0)
{
    init(cache, name);
}

OrthoCPUProducer::OrthoCPUProducer() :
    TileProducer("OrthoCPUProducer", "CreateOrthoCPUTile"), channels(0), tileSize(0), border(2), maxLevel(-1), dxt(false),
    header(0), fileData(NULL), fileSize(0), storage(NULL), tileCount(0)
{
}

void OrthoCPUProducer::init(ptr<TileCache> cache, const char *name)
{
    /* a CPU producer in the reference (worker threads); the decode is device work here */
    TileProducer::init(cache, true);
    this->name = name;
    storage = dynamic_cast<GPUTileStorage *>(cache->getStorage().get());
    if (storage == NULL || (storage->getInternalFormat() != RGBA8 && storage->getInternalFormat() != RGB8)) {
        if (Logger::ERROR_LOGGER != NULL) {
            Logger::ERROR_LOGGER->log("ORTHO", "OrthoCPUProducer needs a cpuByteTileStorage (device byte pool)");
        }
        throw std::invalid_argument("OrthoCPUProducer: bad tile storage");
    }
    context = storage->getContext();
    channels = storage->getComponents();
    try {
        load(name);
    } catch (...) {
        /* a constructor that throws runs no destructor: leave nothing behind */
        if (fileData != NULL) {
            munmap(const_cast<unsigned char *>(fileData), fileSize);
            fileData = NULL;
        }
        throw;
    }
    context->addSource(this);
}

void OrthoCPUProducer::load(const char *name)
{
    dxt = false;
    border = 2;
    header = 0;
    if (strlen(name) == 0) {
        /* no file: all-zero tiles of levels 0..1 (OrthoCPUProducer.cpp:73-77) */
        maxLevel = 1;
        tileSize = 0;
        return;
    }
    maxLevel = -1;
    tileSize = storage->getTileSize() - 4;
    const int fd = open(name, O_RDONLY);
    struct stat st;
    if (fd < 0 || fstat(fd, &st) != 0 || st.st_size < 28) {
        if (fd >= 0) close(fd);
        if (Logger::ERROR_LOGGER != NULL) {
            Logger::ERROR_LOGGER->log("ORTHO", "Cannot open file '" + std::string(name) + "'");
        }
        return;     /* like the reference: maxLevel = -1, the producer has no tile */
    }
    void *map = mmap(NULL, (size_t) st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (map == MAP_FAILED) {
        throw DeviceError(PL_ERR_IO, "cannot map '" + std::string(name) + "'");
    }
    fileData = static_cast<const unsigned char *>(map);
    fileSize = (size_t) st.st_size;
    int head[7];
    memcpy(head, fileData, sizeof(head));
    maxLevel = head[0];
    tileSize = head[1];
    channels = head[2];
    dxt = (head[6] & 1) != 0;
    border = (head[6] & 2) != 0 ? 0 : 2;
    if (maxLevel < 0 || maxLevel > 14) {
        throw DeviceError(PL_ERR_CORRUPT, "'" + std::string(name) + "': bad header");
    }
    const int ntiles = ((1 << (maxLevel * 2 + 2)) - 1) / 3;
    header = 7 * sizeof(int) + 2 * ntiles * sizeof(long long);
    if (fileSize < header) {
        throw DeviceError(PL_ERR_CORRUPT, "'" + std::string(name) + "': offset table is truncated");
    }
    offsets.assign((size_t) ntiles * 2, 0);
    memcpy(&offsets[0], fileData + 28, sizeof(long long) * ntiles * 2);
    if (dxt) {
        throw DeviceError(PL_ERR_CORRUPT, "'" + std::string(name) + "': DXT-compressed ortho files are not supported");
    }
    /* the asserts of OrthoCPUProducer.cpp:178-180 */
    if (storage->getComponents() != channels && !(storage->getInternalFormat() == RGBA8 && channels <= 4)) {
        throw std::invalid_argument("OrthoCPUProducer: the storage must have the file's channel count");
    }
    if (storage->getTileSize() != tileSize + 2 * border) {
        throw std::invalid_argument("OrthoCPUProducer: the storage tile size must be the file's tile size + 2 * border");
    }
}

OrthoCPUProducer::~OrthoCPUProducer()
{
    if (context != NULL) {
        context->removeSource(this);
    }
    if (fileData != NULL) {
        munmap(const_cast<unsigned char *>(fileData), fileSize);
    }
}

int OrthoCPUProducer::getBorder()
{
    return border;
}

bool OrthoCPUProducer::hasTile(int level, int tx, int ty)
{
    (void) tx;
    (void) ty;
    return level <= maxLevel;
}

bool OrthoCPUProducer::isCompressed()
{
    return dxt;
}

int OrthoCPUProducer::getTileId(int level, int tx, int ty)
{
    return tx + ty * (1 << level) + ((1 << (2 * level)) - 1) / 3;
}

bool OrthoCPUProducer::doCreateTile(int level, int tx, int ty, TileStorage::Slot *data)
{
    if (Logger::DEBUG_LOGGER != NULL) {
        Logger::DEBUG_LOGGER->logf("ORTHO", "CPU tile %d %d %d %d", getId(), level, tx, ty);
    }
    GPUTileStorage::GPUSlot *slot = dynamic_cast<GPUTileStorage::GPUSlot *>(data);
    assert(slot != NULL);
    assert(name.empty() || level <= maxLevel);
    Job j = { getTileId(level, tx, ty), slot->l };
    pending.push_back(j);
    ++tileCount;
    return true;
}

void OrthoCPUProducer::endCreateTile()
{
    TileProducer::endCreateTile();
    if (!context->inBatch()) {
        context->flush();
    }
}

void OrthoCPUProducer::flushBatch()
{
    if (pending.empty()) {
        return;
    }
    std::vector<Job> jobs;
    jobs.swap(pending);
    pl_pool *pool = storage->getPool();
    if (name.empty()) {
        const int w = storage->getTileSize();
        std::vector<unsigned char> zeros((size_t) w * w * 4, 0);
        for (size_t i = 0; i < jobs.size(); ++i) {
            DeviceContext::check(pl_pool_upload(pool, jobs[i].slot, &zeros[0], zeros.size()));
        }
        return;
    }
    std::vector<uint64_t> offs(jobs.size());
    std::vector<uint32_t> sizes(jobs.size());
    std::vector<int32_t> slots(jobs.size());
    for (size_t i = 0; i < jobs.size(); ++i) {
        const int id = jobs[i].tileid;
        if (id < 0 || (size_t) (2 * id + 1) >= offsets.size()) {
            throw DeviceError(PL_ERR_CORRUPT, "'" + name + "': tile id out of range");
        }
        const long long a = offsets[2 * id], b = offsets[2 * id + 1];
        if (a < 0 || b < a || (uint64_t) header + (uint64_t) b > fileSize) {
            throw DeviceError(PL_ERR_CORRUPT, "'" + name + "': blob outside the file");
        }
        offs[i] = (uint64_t) header + (uint64_t) a;
        sizes[i] = (uint32_t) (b - a);
        slots[i] = jobs[i].slot;
    }
    int fileChannels = 0;
    DeviceContext::check(pl_ortho_decode_batch(context->handle(), pool, (int) jobs.size(), fileData, &offs[0], &sizes[0],
                                               &slots[0], &fileChannels));
    if (fileChannels != channels) {
        throw DeviceError(PL_ERR_CORRUPT, "'" + name + "': the blobs' channel count differs from the header's");
    }
}

}  // namespace proland
