#include "proland/producer/DeviceContext.h"

#include <algorithm>
#include <map>
#include <mutex>

#include "ork/BatchScheduler.h"

namespace proland
{

static std::mutex g_mutex;
static std::map<int, ptr<DeviceContext> > g_contexts;
static int g_current = 0;

void DeviceContext::check(int status)
{
    if (status != PL_OK) {
        const char *msg = pl_last_error();
        if (Logger::ERROR_LOGGER != NULL) {
            Logger::ERROR_LOGGER->log("DEVICE", msg ? msg : "error");
        }
        throw DeviceError(status, msg ? msg : "error");
    }
}

ptr<DeviceContext> DeviceContext::get(int device)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (device < 0) {
        device = g_current;
    }
    std::map<int, ptr<DeviceContext> >::iterator i = g_contexts.find(device);
    if (i != g_contexts.end()) {
        return i->second;
    }
    pl_ctx *ctx = NULL;
    check(pl_ctx_create(device, &ctx));
    ptr<DeviceContext> c = new DeviceContext(device, ctx);
    g_contexts[device] = c;
    BatchScheduler::setWaveHook(&DeviceContext::waveHook);
    return c;
}

void DeviceContext::setCurrentDevice(int device)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    g_current = device;
}

int DeviceContext::getCurrentDevice()
{
    std::lock_guard<std::mutex> lock(g_mutex);
    return g_current;
}

void DeviceContext::shutdown()
{
    std::lock_guard<std::mutex> lock(g_mutex);
    g_contexts.clear();
}

DeviceContext::DeviceContext(int device, pl_ctx *ctx) :
    Object("DeviceContext"), device(device), ctx(ctx), noiseWidth(0), orthoNoiseWidth(0), depth(0)
{
}

DeviceContext::~DeviceContext()
{
    pl_ctx_destroy(ctx);
}

void DeviceContext::ensureNoise(int tileWidth)
{
    /* the library keeps one table per width (demNoiseFactory caches one texture per width,
     * ElevationProducer.cpp:135): producers of different tile sizes can share this context */
    if (noiseWidth != tileWidth) {
        check(pl_noise_init(ctx, tileWidth, NULL));
        noiseWidth = tileWidth;
    }
}

void DeviceContext::ensureOrthoNoise(int tileWidth)
{
    if (orthoNoiseWidth != tileWidth) {
        flush();
        check(pl_ortho_noise_init(ctx, tileWidth, NULL));
        orthoNoiseWidth = tileWidth;
    }
}

BatchSourceRegistration::BatchSourceRegistration(DeviceContext *context, BatchSource *source) :
    context(context), source(source), committed(false)
{
    context->addSource(source);
}

BatchSourceRegistration::~BatchSourceRegistration()
{
    if (!committed) {
        context->removeSource(source);
    }
}

void DeviceContext::addSource(BatchSource *s)
{
    sources.push_back(s);
}

void DeviceContext::removeSource(BatchSource *s)
{
    sources.erase(std::remove(sources.begin(), sources.end(), s), sources.end());
}

void DeviceContext::beginBatch()
{
    ++depth;
}

void DeviceContext::endBatch()
{
    if (--depth == 0) {
        flush();
    }
}

void DeviceContext::flush()
{
    /* sources are registered in creation order: a producer is created after the producers it reads
     * (the resource loader resolves `residuals=` / `elevations=` first), so this is a valid launch
     * order even for tiles queued in the same batch */
    for (size_t i = 0; i < sources.size(); ++i) {
        sources[i]->flushBatch();
    }
}

void DeviceContext::waveHook(bool begin)
{
    std::vector<ptr<DeviceContext> > all;
    {
        std::lock_guard<std::mutex> lock(g_mutex);
        for (std::map<int, ptr<DeviceContext> >::iterator i = g_contexts.begin(); i != g_contexts.end(); ++i) {
            all.push_back(i->second);
        }
    }
    for (size_t i = 0; i < all.size(); ++i) {
        if (begin) {
            all[i]->beginBatch();
        } else {
            all[i]->endBatch();
        }
    }
}

void DeviceContext::sync()
{
    check(pl_sync(ctx));
}

unsigned long long DeviceContext::getLaunchCount()
{
    return pl_ctx_launch_count(ctx);
}

}  // namespace proland
