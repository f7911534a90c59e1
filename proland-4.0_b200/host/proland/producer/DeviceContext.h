/*
 * DeviceContext -- what the GL context + FBO is to the reference's GPU producers
 * (ElevationProducer.cpp:136-155, NormalProducer.cpp:48-64): the per-GPU state all
 * producers of a scene share.  It wraps one pl_ctx of the C ABI
 * (include/proland_b200.h) and collects the tiles the producers enqueue in
 * doCreateTile into batches: one kernel launch per producer per flush instead of
 * one draw call per tile.
 */
#ifndef PROLAND_B200_DEVICE_CONTEXT_H
#define PROLAND_B200_DEVICE_CONTEXT_H

#include <stdexcept>
#include <string>
#include <vector>

#include "ork/ork_lite.h"
#include "proland_b200.h"

using namespace ork;

namespace proland
{

/* a C-ABI status other than PL_OK; what() carries pl_last_error() */
class DeviceError : public std::runtime_error
{
public:
    DeviceError(int code, const std::string &what) : std::runtime_error(what), code(code) {}
    int code;
};

/* implemented by producers that queue device work */
class BatchSource
{
public:
    virtual ~BatchSource() {}
    /* launch everything queued since the last flush */
    virtual void flushBatch() = 0;
};

class DeviceContext;

/* Registers a producer with its context for the length of an init(); unless commit() is reached the
 * registration is undone -- a constructor that throws runs no destructor, and the context must not keep a
 * pointer to the dead object. */
class BatchSourceRegistration
{
public:
    BatchSourceRegistration(DeviceContext *context, BatchSource *source);
    ~BatchSourceRegistration();
    void commit() { committed = true; }

private:
    DeviceContext *context;
    BatchSource *source;
    bool committed;
};

class DeviceContext : public Object
{
public:
    /* the context of a CUDA device (created on first use); device < 0: the current one.
     * Throws DeviceError(PL_ERR_NO_DEVICE) when there is no GPU: there is no CPU fallback. */
    static ptr<DeviceContext> get(int device = -1);
    static void setCurrentDevice(int device);
    static int getCurrentDevice();
    /* destroys the cached contexts (tests) */
    static void shutdown();

    virtual ~DeviceContext();

    pl_ctx *handle() { return ctx; }
    int getDevice() const { return device; }
    /* createDemNoise is cached per tile width in the reference (demNoiseFactory,
     * ElevationProducer.cpp:135); one width per context here */
    void ensureNoise(int tileWidth);
    /* createOrthoNoise, cached per tile width (orthoNoiseFactory, OrthoProducer.cpp:120) */
    void ensureOrthoNoise(int tileWidth);

    void addSource(BatchSource *s);
    void removeSource(BatchSource *s);
    /* between beginBatch and endBatch producers only queue; endBatch launches */
    void beginBatch();
    void endBatch();
    bool inBatch() const { return depth > 0; }
    void flush();
    void sync();
    unsigned long long getLaunchCount();

    static void check(int status);
    /* BatchScheduler's wave hook: begin / end a batch on every live context */
    static void waveHook(bool begin);

private:
    DeviceContext(int device, pl_ctx *ctx);
    int device;
    pl_ctx *ctx;
    int noiseWidth, orthoNoiseWidth;
    int depth;
    std::vector<BatchSource *> sources;
};

}  // namespace proland

#endif
