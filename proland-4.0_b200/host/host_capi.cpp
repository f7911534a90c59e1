/*
 * host_capi.cpp -- flat C entry points over the C++ host layer (namespace
 * proland), for the Python test-suite, bench.py and tools.  Not the drop-in
 * boundary (that is include/proland_b200.h); a C++ application links the classes
 * directly.  Declared in host/proland_host.h.
 */
#include "proland_host.h"

#include <cstring>
#include <string>
#include <vector>

#include "ork/BatchScheduler.h"
#include "proland/dem/ElevationProducer.h"
#include "proland/dem/NormalProducer.h"
#include "proland/dem/ResidualProducer.h"
#include "proland/ortho/OrthoCPUProducer.h"
#include "proland/ortho/OrthoProducer.h"
#include "proland/preprocess/terrain/Preprocess.h"
#include "proland/producer/TileCache.h"
#include "proland/producer/TileProducer.h"
#include "proland/resource/ResourceManager.h"
#include "proland/terrain/TerrainQuad.h"
#include "proland/terrain/TileSampler.h"
#include "proland/terrain/TileSamplerZ.h"

using namespace proland;

static thread_local std::string g_error;

#define PLH_TRY try {
#define PLH_CATCH(ret)                                        \
    } catch (const DeviceError &e) {                          \
        g_error = e.what();                                   \
        return ret;                                           \
    } catch (const std::exception &e) {                       \
        g_error = e.what();                                   \
        return ret;                                           \
    }

namespace
{

/* A producer without device work: tiles depend on their parent like elevation tiles do; doCreateTile
 * only records what it was asked.  Lets the cache / task / scheduler logic run on a CPU-only box. */
class RecordingProducer : public TileProducer
{
public:
    struct Call { int level, tx, ty, slot; };
    std::vector<Call> calls;
    int maxLevel;
    int begins, ends;

    RecordingProducer(ptr<TileCache> cache, int maxLevel) :
        TileProducer("RecordingProducer", "CreateRecordedTile", cache, false), maxLevel(maxLevel), begins(0), ends(0)
    {
    }
    virtual bool hasTile(int level, int tx, int ty) { (void) tx; (void) ty; return level <= maxLevel; }

protected:
    virtual ptr<Task> startCreateTile(int level, int tx, int ty, unsigned int deadline, ptr<Task> task, ptr<TaskGraph> owner)
    {
        ptr<TaskGraph> result = owner == NULL ? createTaskGraph(task) : owner;
        if (level > 0) {
            TileCache::Tile *t = getTile(level - 1, tx / 2, ty / 2, deadline);
            if (t == NULL) {
                cacheFull("RecordingProducer");
            }
            result->addTask(t->task);
            result->addDependency(task, t->task);
        }
        return result;
    }
    virtual void beginCreateTile() { ++begins; }
    virtual bool doCreateTile(int level, int tx, int ty, TileStorage::Slot *data)
    {
        Call c = { level, tx, ty, slotIndex(data) };
        calls.push_back(c);
        return true;
    }
    virtual void endCreateTile() { ++ends; }
    virtual void stopCreateTile(int level, int tx, int ty)
    {
        if (level > 0) {
            TileCache::Tile *t = findTile(level - 1, tx / 2, ty / 2);
            if (t != NULL) putTile(t);
        }
    }

private:
    std::vector<TileStorage::Slot *> known;
    int slotIndex(TileStorage::Slot *s)
    {
        for (size_t i = 0; i < known.size(); ++i) if (known[i] == s) return (int) i;
        known.push_back(s);
        return (int) known.size() - 1;
    }
};

class PlainStorage : public TileStorage
{
public:
    PlainStorage(int tileSize, int capacity) : TileStorage(tileSize, capacity)
    {
        for (int i = 0; i < capacity; ++i) freeSlots.push_back(new Slot(this));
    }
};

struct TestScene
{
    ptr<BatchScheduler> scheduler;
    ptr<TileCache> cache;
    ptr<RecordingProducer> producer;
};

static Logger g_debug("DEBUG");

}  // namespace

extern "C" {

const char *plh_last_error(void) { return g_error.c_str(); }

void *plh_open(const char *xml, const char *data_dir, int device)
{
    PLH_TRY
    ResourceManager *m = new ResourceManager(xml, data_dir ? data_dir : ".", device);
    m->acquire();
    return m;
    PLH_CATCH(NULL)
}

void plh_close(void *mgr)
{
    if (mgr != NULL) {
        ResourceManager *m = static_cast<ResourceManager *>(mgr);
        m->close();
        m->release();
    }
}

void plh_shutdown(void) { DeviceContext::shutdown(); }

static Object *load(void *mgr, const char *name)
{
    return static_cast<ResourceManager *>(mgr)->loadResource(name).get();
}

void *plh_producer(void *mgr, const char *name)
{
    PLH_TRY
    TileProducer *p = dynamic_cast<TileProducer *>(load(mgr, name));
    if (p == NULL) g_error = std::string("'") + name + "' is not a TileProducer";
    return p;
    PLH_CATCH(NULL)
}

void *plh_cache(void *mgr, const char *name)
{
    PLH_TRY
    TileCache *c = dynamic_cast<TileCache *>(load(mgr, name));
    if (c == NULL) g_error = std::string("'") + name + "' is not a TileCache";
    return c;
    PLH_CATCH(NULL)
}

void *plh_scheduler(void *mgr, const char *name)
{
    PLH_TRY
    BatchScheduler *s = dynamic_cast<BatchScheduler *>(load(mgr, name));
    if (s == NULL) g_error = std::string("'") + name + "' is not a scheduler";
    return s;
    PLH_CATCH(NULL)
}

void *plh_producer_cache(void *prod) { return static_cast<TileProducer *>(prod)->getCache().get(); }
void *plh_cache_scheduler(void *cache) { return dynamic_cast<BatchScheduler *>(static_cast<TileCache *>(cache)->getScheduler().get()); }

void plh_set_root_quad_size(void *prod, float size) { static_cast<TileProducer *>(prod)->setRootQuadSize(size); }

int plh_producer_info(void *prod, int out[6])
{
    TileProducer *p = static_cast<TileProducer *>(prod);
    out[0] = p->getId();
    out[1] = p->getBorder();
    out[2] = p->isGpuProducer() ? 1 : 0;
    out[3] = p->getCache()->getStorage()->getTileSize();
    std::vector<ptr<TileProducer> > refs;
    p->getReferencedProducers(refs);
    out[4] = (int) refs.size();
    out[5] = (int) (p->getRootQuadSize());
    return 0;
}

const char *plh_producer_type(void *prod) { return static_cast<TileProducer *>(prod)->getClass(); }
const char *plh_producer_task_type(void *prod) { return static_cast<TileProducer *>(prod)->getTaskType(); }


int plh_has_tile(void *prod, int level, int tx, int ty) { return static_cast<TileProducer *>(prod)->hasTile(level, tx, ty) ? 1 : 0; }
int plh_has_children(void *prod, int level, int tx, int ty) { return static_cast<TileProducer *>(prod)->hasChildren(level, tx, ty) ? 1 : 0; }

void *plh_get_tile(void *prod, int level, int tx, int ty, unsigned int deadline)
{
    PLH_TRY
    TileCache::Tile *t = static_cast<TileProducer *>(prod)->getTile(level, tx, ty, deadline);
    if (t == NULL) g_error = "Insufficient tile cache size";
    return t;
    PLH_CATCH(NULL)
}

void *plh_find_tile(void *prod, int level, int tx, int ty, int include_cache, int done)
{
    return static_cast<TileProducer *>(prod)->findTile(level, tx, ty, include_cache != 0, done != 0);
}

int plh_put_tile(void *prod, void *tile)
{
    PLH_TRY
    static_cast<TileProducer *>(prod)->putTile(static_cast<TileCache::Tile *>(tile));
    return 0;
    PLH_CATCH(-1)
}

int plh_prefetch_tile(void *prod, int level, int tx, int ty)
{
    PLH_TRY
    return static_cast<TileProducer *>(prod)->prefetchTile(level, tx, ty) ? 1 : 0;
    PLH_CATCH(-1)
}

void plh_invalidate_tiles(void *prod) { static_cast<TileProducer *>(prod)->invalidateTiles(); }
void plh_invalidate_tile(void *prod, int level, int tx, int ty) { static_cast<TileProducer *>(prod)->invalidateTile(level, tx, ty); }

int plh_tile_done(void *tile) { return static_cast<TileCache::Tile *>(tile)->task->isDone() ? 1 : 0; }

int plh_tile_slot(void *tile)
{
    TileCache::Tile *t = static_cast<TileCache::Tile *>(tile);
    GPUTileStorage::GPUSlot *s = dynamic_cast<GPUTileStorage::GPUSlot *>(t->getData(false));
    return s == NULL ? -1 : s->l;
}

int plh_tile_download(void *tile, void *buf, size_t bytes)
{
    PLH_TRY
    TileCache::Tile *t = static_cast<TileCache::Tile *>(tile);
    GPUTileStorage::GPUSlot *s = dynamic_cast<GPUTileStorage::GPUSlot *>(t->getData());
    if (s == NULL) {
        g_error = "tile is not done or not device-resident";
        return -1;
    }
    s->getImage(buf, bytes);
    return 0;
    PLH_CATCH(-1)
}

int plh_tile_minmax(void *prod, void *tile, float out[2])
{
    PLH_TRY
    ElevationProducer *e = dynamic_cast<ElevationProducer *>(static_cast<TileProducer *>(prod));
    if (e == NULL) {
        g_error = "not an ElevationProducer";
        return -1;
    }
    e->getTileMinMax(static_cast<TileCache::Tile *>(tile), &out[0], &out[1]);
    return 0;
    PLH_CATCH(-1)
}

/* what TileSampler::update does with the tiles it needs: one graph of their tasks, one Scheduler::run */
int plh_run(void *scheduler, void **tiles, int n)
{
    PLH_TRY
    ptr<TaskGraph> g = new TaskGraph();
    for (int i = 0; i < n; ++i) {
        g->addTask(static_cast<TileCache::Tile *>(tiles[i])->task);
    }
    static_cast<BatchScheduler *>(scheduler)->run(g);
    return 0;
    PLH_CATCH(-1)
}

int plh_cache_stats(void *cache, int out[6])
{
    TileCache *c = static_cast<TileCache *>(cache);
    out[0] = c->getUsedTiles();
    out[1] = c->getUnusedTiles();
    out[2] = c->getStorage()->getCapacity();
    out[3] = c->getStorage()->getFreeSlots();
    out[4] = c->getQueries();
    out[5] = c->getMisses();
    return 0;
}

int plh_scheduler_stats(void *scheduler, unsigned long long out[4])
{
    BatchScheduler *s = static_cast<BatchScheduler *>(scheduler);
    out[0] = s->getFrame();
    out[1] = s->getWaveCount();
    out[2] = s->getTaskCount();
    out[3] = (unsigned long long) s->getQueuedPrefetchCount();
    return 0;
}

int plh_producer_counts(void *prod, unsigned long long out[2])
{
    TileProducer *p = static_cast<TileProducer *>(prod);
    out[0] = out[1] = 0;
    if (ElevationProducer *e = dynamic_cast<ElevationProducer *>(p)) { out[0] = e->getTileCount(); out[1] = e->getBatchCount(); }
    else if (NormalProducer *n = dynamic_cast<NormalProducer *>(p)) { out[0] = n->getTileCount(); out[1] = n->getBatchCount(); }
    else if (OrthoProducer *o = dynamic_cast<OrthoProducer *>(p)) { out[0] = o->getTileCount(); out[1] = o->getBatchCount(); }
    else if (OrthoCPUProducer *c = dynamic_cast<OrthoCPUProducer *>(p)) { out[0] = c->getTileCount(); }
    else if (ResidualProducer *r = dynamic_cast<ResidualProducer *>(p)) { out[0] = r->getTileCount(); }
    else return -1;
    return 0;
}

int plh_residual_info(void *prod, int out[3])
{
    ResidualProducer *r = dynamic_cast<ResidualProducer *>(static_cast<TileProducer *>(prod));
    if (r == NULL) return -1;
    out[0] = r->getMinLevel();
    out[1] = r->getMaxLevel();
    out[2] = r->getDeltaLevel();
    return 0;
}

int plh_residual_tile_id(void *prod, int level, int tx, int ty)
{
    ResidualProducer *r = dynamic_cast<ResidualProducer *>(static_cast<TileProducer *>(prod));
    return r == NULL ? -1 : r->getTileId(level, tx, ty);
}

int plh_residual_tile_size(void *prod, int level)
{
    ResidualProducer *r = dynamic_cast<ResidualProducer *>(static_cast<TileProducer *>(prod));
    return r == NULL ? -1 : r->getTileSize(level);
}

unsigned long long plh_device_launches(int device)
{
    PLH_TRY
    return DeviceContext::get(device)->getLaunchCount();
    PLH_CATCH(0)
}

int plh_device_sync(int device)
{
    PLH_TRY
    DeviceContext::get(device)->sync();
    return 0;
    PLH_CATCH(-1)
}

void plh_debug_log(int on, int echo)
{
    g_debug.setEcho(echo != 0);
    Logger::DEBUG_LOGGER = on ? &g_debug : NULL;
}
unsigned long plh_debug_log_lines(void) { return g_debug.getLineCount(); }
void plh_quiet_errors(int quiet) { if (Logger::ERROR_LOGGER) Logger::ERROR_LOGGER->setEcho(!quiet); if (Logger::WARNING_LOGGER) Logger::WARNING_LOGGER->setEcho(!quiet); }

int plh_upsample_variant(const char *prog, int out[2])
{
    bool slope, noclamp;
    const bool ok = ResourceManager::upsampleVariant(prog, &slope, &noclamp);
    out[0] = slope;
    out[1] = noclamp;
    return ok ? 0 : -1;
}

/* ---------------------------------------------------------------- terrain quadtree + samplers */

void *plh_terrain_create(float size, float zmin, float zmax, float split_factor, int max_level)
{
    PLH_TRY
    TerrainNode *n = new TerrainNode(size, zmin, zmax, split_factor, max_level);
    n->acquire();
    return n;
    PLH_CATCH(NULL)
}

void plh_terrain_destroy(void *node) { if (node) static_cast<TerrainNode *>(node)->release(); }

float plh_split_distance(float split_factor, float viewport_width, float fov_radians)
{
    return TerrainNode::splitDistance(split_factor, viewport_width, fov_radians);
}

int plh_terrain_update(void *node, double x, double y, double z, float split_dist, float dist_factor)
{
    PLH_TRY
    TerrainNode *n = static_cast<TerrainNode *>(node);
    n->update(x, y, z, split_dist, dist_factor);
    return n->root->getSize();
    PLH_CATCH(-1)
}

static void list_quads(TerrainQuad *q, int *out, int max_quads, int *n)
{
    if (*n < max_quads) {
        out[4 * *n] = q->level;
        out[4 * *n + 1] = q->tx;
        out[4 * *n + 2] = q->ty;
        out[4 * *n + 3] = q->isLeaf() ? 1 : 0;
    }
    ++*n;
    if (!q->isLeaf()) {
        for (int i = 0; i < 4; ++i) list_quads(q->children[i].get(), out, max_quads, n);
    }
}

/* the quadtree in pre-order (children 0..3) as (level, tx, ty, leaf) quadruples; returns the quad count */
int plh_terrain_quads(void *node, int *out, int max_quads)
{
    int n = 0;
    list_quads(static_cast<TerrainNode *>(node)->root.get(), out, max_quads, &n);
    return n;
}

void *plh_sampler_create(const char *name, void *prod, int async, int store_parent)
{
    PLH_TRY
    TileSampler *s = new TileSampler(name, static_cast<TileProducer *>(prod));
    s->setStoreParent(store_parent != 0);
    s->setAsynchronous(async != 0);
    s->acquire();
    return s;
    PLH_CATCH(NULL)
}

/* a TileSamplerZ (<tileSamplerZ sampler= producer=>): also feeds TerrainQuad::zmin / zmax and the ground height under the
 * camera from the z ranges of the elevation tiles it holds */
void *plh_sampler_z_create(const char *name, void *prod, int async, int store_parent)
{
    PLH_TRY
    TileSamplerZ *s = new TileSamplerZ(name, static_cast<TileProducer *>(prod));
    s->setStoreParent(store_parent != 0);
    s->setAsynchronous(async != 0);
    s->acquire();
    return static_cast<TileSampler *>(s);
    PLH_CATCH(NULL)
}

int plh_sampler_z_counts(void *sampler, unsigned long long out[3])
{
    TileSamplerZ *z = dynamic_cast<TileSamplerZ *>(static_cast<TileSampler *>(sampler));
    if (z == NULL) return -1;
    z->getCounts(out);
    return 0;
}

/* TerrainNode::groundHeightAtCamera and nextGroundHeightAtCamera; set >= 0: overwrite both first (tests) */
void plh_ground_height(float out[2], int reset)
{
    if (reset) TerrainNode::groundHeightAtCamera = TerrainNode::nextGroundHeightAtCamera = 0.0f;
    out[0] = TerrainNode::groundHeightAtCamera;
    out[1] = TerrainNode::nextGroundHeightAtCamera;
}

static void list_quads_z(TerrainQuad *q, float *out, int max_quads, int *n)
{
    if (*n < max_quads) {
        float *o = out + 5 * (*n);
        o[0] = (float) q->level; o[1] = (float) q->tx; o[2] = (float) q->ty; o[3] = q->zmin; o[4] = q->zmax;
    }
    ++*n;
    if (!q->isLeaf())
        for (int i = 0; i < 4; ++i) list_quads_z(q->children[i].get(), out, max_quads, n);
}

/* pre-order (level, tx, ty, zmin, zmax) as floats */
int plh_terrain_quads_z(void *node, float *out, int max_quads)
{
    int n = 0;
    list_quads_z(static_cast<TerrainNode *>(node)->root.get(), out, max_quads, &n);
    return n;
}

void plh_sampler_destroy(void *sampler)
{
    if (sampler) {
        static_cast<TileSampler *>(sampler)->release();
        static_cast<Object *>(static_cast<TileSampler *>(sampler))->release();
    }
}

int plh_sampler_tile_count(void *sampler) { return static_cast<TileSampler *>(sampler)->getTileCount(); }

/* One frame: every sampler updates against the terrain's quadtree (putTiles / getTiles / prefetch), the
 * tasks they return form one graph, the scheduler runs it.  Returns the number of tasks the graph held. */
int plh_frame_update(void *scheduler, void *node, void **samplers, int n)
{
    PLH_TRY
    ptr<TaskGraph> frame = new TaskGraph();
    int tasks = 0;
    static unsigned int frameNumber = 0;      /* SceneManager::getFrameNumber() */
    ++frameNumber;
    for (int i = 0; i < n; ++i) {
        ptr<TaskGraph> g = static_cast<TileSampler *>(samplers[i])->update(static_cast<TerrainNode *>(node)->root, frameNumber);
        TaskGraph::TaskIterator it = g->getAllTasks();
        while (it.hasNext()) {
            frame->addTask(it.next());
            ++tasks;
        }
    }
    static_cast<BatchScheduler *>(scheduler)->run(frame);
    return tasks;
    PLH_CATCH(-1)
}

/* ---------------------------------------------------------------- CPU-only test double */

int plh_preprocess_dem(const float *src, int src_w, int src_h, int min_tile_size, int tile_size, int max_level, const char *dst_folder,
                       float residual_scale, int spherical)
{
    PLH_TRY
    ArrayInputMap map(src, src_w, src_h);
    if (spherical) preprocessSphericalDem(&map, min_tile_size, tile_size, max_level, dst_folder, "", residual_scale);
    else preprocessDem(&map, min_tile_size, tile_size, max_level, dst_folder, "", residual_scale);
    return 0;
    PLH_CATCH(-1)
}

void *plh_test_scene(int capacity, int tile_size, int max_level, int prefetch_rate, int prefetch_queue)
{
    PLH_TRY
    TestScene *s = new TestScene();
    s->scheduler = new BatchScheduler(prefetch_rate, prefetch_queue);
    s->cache = new TileCache(new PlainStorage(tile_size, capacity), "test", s->scheduler);
    s->producer = new RecordingProducer(s->cache, max_level);
    return s;
    PLH_CATCH(NULL)
}

void plh_test_scene_close(void *scene)
{
    TestScene *s = static_cast<TestScene *>(scene);
    s->scheduler->clear();
    s->producer = NULL;
    s->cache = NULL;
    s->scheduler = NULL;
    delete s;
}

void *plh_test_producer(void *scene) { return static_cast<TestScene *>(scene)->producer.get(); }
void *plh_test_cache(void *scene) { return static_cast<TestScene *>(scene)->cache.get(); }
void *plh_test_scheduler(void *scene) { return static_cast<TestScene *>(scene)->scheduler.get(); }

/* the doCreateTile calls so far as (level, tx, ty, slot) quadruples; returns their number */
int plh_test_calls(void *scene, int *out, int max_calls)
{
    RecordingProducer *p = static_cast<TestScene *>(scene)->producer.get();
    const int n = (int) p->calls.size();
    for (int i = 0; i < n && i < max_calls; ++i) {
        out[4 * i] = p->calls[i].level;
        out[4 * i + 1] = p->calls[i].tx;
        out[4 * i + 2] = p->calls[i].ty;
        out[4 * i + 3] = p->calls[i].slot;
    }
    return n;
}

int plh_test_begin_end(void *scene, int out[2])
{
    RecordingProducer *p = static_cast<TestScene *>(scene)->producer.get();
    out[0] = p->begins;
    out[1] = p->ends;
    return 0;
}

}  // extern "C"
