/*
 * CPUTileStorage<float> -- the residual tile pool.
 *
 * In the reference (producer/CPUTileStorage.h:43-155) residual tiles live in host
 * arrays (`CPUSlot::data`) that ElevationProducer::doCreateTile copies, window by
 * window, into a texture for every tile (ElevationProducer.cpp:326-339).  Here the
 * decoded residuals never leave the device: the storage is a float pool in HBM
 * (PL_POOL_RESID_F32) written by the decode kernels and read in place by the
 * elevation kernel.  A CPUSlot therefore has no `data` pointer; getData() copies
 * the tile out for hosts and tests.
 */
#ifndef PROLAND_B200_CPU_TILE_STORAGE_H
#define PROLAND_B200_CPU_TILE_STORAGE_H

#include "proland/producer/GPUTileStorage.h"

namespace proland
{

template <class T>
class CPUTileStorage;

template <>
PROLAND_API class CPUTileStorage<float> : public GPUTileStorage
{
public:
    typedef GPUTileStorage::GPUSlot CPUSlot;

    CPUTileStorage(int tileSize, int channels, int capacity, ptr<DeviceContext> context = NULL) :
        GPUTileStorage(tileSize, capacity, R32F, NEAREST, NEAREST, context), channels(channels)
    {
        if (channels != 1) {
            throw std::invalid_argument("cpuFloatTileStorage: residual tiles have one channel");
        }
    }
    int getChannels() { return channels; }
    /* tileSize * tileSize floats, row stride tileSize: the reference's CPUSlot::data layout */
    static void getData(CPUSlot *slot, float *out)
    {
        const int w = slot->getOwner()->getTileSize();
        slot->getImage(out, (size_t) w * w * sizeof(float));
    }

private:
    int channels;
};

/* CPUTileStorage<unsigned char> (`cpuByteTileStorage`): the ortho residual tile pool.  In the reference
 * OrthoProducer::doCreateTile uploads CPUSlot::data into residualTexture for every tile
 * (ortho/OrthoProducer.cpp:296-318); here the decoded bytes stay on the device in an RGBA8 pool
 * (PL_POOL_ORTHO_UN8x4 / _NORM_UN8x4; a tile with fewer channels uses the first bytes of each texel). */
template <>
PROLAND_API class CPUTileStorage<unsigned char> : public GPUTileStorage
{
public:
    typedef GPUTileStorage::GPUSlot CPUSlot;

    CPUTileStorage(int tileSize, int channels, int capacity, ptr<DeviceContext> context = NULL) :
        GPUTileStorage(tileSize, capacity, channels == 4 ? RGBA8 : RGB8, NEAREST, NEAREST, context), channels(channels)
    {
        if (channels < 1 || channels > 4) {
            throw std::invalid_argument("cpuByteTileStorage: 1 to 4 channels");
        }
    }
    int getChannels() { return channels; }
    virtual int getComponents() const { return channels; }

private:
    int channels;
};

}  // namespace proland

#endif
