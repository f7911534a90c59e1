/*
 * GPUTileStorage -- the device-resident tile pool.
 *
 * Replaces the reference's GPUTileStorage (producer/GPUTileStorage.h:52-250,
 * GPUTileStorage.cpp:119-199): there a slot is a layer `l` of a GL 2D texture
 * array and the only writers are copyPixels / setSubImage; here a slot is index
 * `l` of one pl_pool slab in HBM (include/proland_b200.h) and the only writers are
 * the batched kernels.  No GL: `internalformat` picks the pool kind, `min`/`mag`
 * are kept because the NormalProducer pass reads elevations through the
 * storage's sampler filter.
 * Not carried over: tileMap (renderer side), mipmaps (ortho storages only).
 */
#ifndef PROLAND_B200_GPU_TILE_STORAGE_H
#define PROLAND_B200_GPU_TILE_STORAGE_H

#include <string>

#include "proland/producer/DeviceContext.h"
#include "proland/producer/TileStorage.h"

namespace proland
{

/* the internal formats the tile-production path uses */
enum TextureInternalFormat { RGB32F, RGBA32F, RG8, RGBA8, R32F, R16I, RGB8 };
enum TextureFilter { NEAREST, LINEAR };

PROLAND_API class GPUTileStorage : public TileStorage
{
public:
    class GPUSlot : public Slot
    {
    public:
        /* the slot's index in the pool (the texture layer in the reference) */
        const int l;
        /* always 0: one pool, where the reference may split into several texture arrays */
        const int index;

        GPUSlot(TileStorage *owner, int index, int l);
        virtual ~GPUSlot();
        int getWidth();
        int getHeight();
        /* the reference's upload path (GPUSlot::setSubImage, GPUTileStorage.cpp:151-171) and its
         * inverse, whole tiles in the reference's pixel layout (pl_pool_upload / pl_pool_download) */
        void setSubImage(const void *pixels, size_t bytes);
        void getImage(void *pixels, size_t bytes);
    };

    GPUTileStorage(int tileSize, int nTiles, TextureInternalFormat internalf, TextureFilter min = NEAREST,
                   TextureFilter mag = NEAREST, ptr<DeviceContext> context = NULL);
    virtual ~GPUTileStorage();

    TextureInternalFormat getInternalFormat() const { return internalf; }
    const char *getInternalFormatName() const;
    virtual int getComponents() const;
    /* LINEAR only when both min and mag are: what a texel fetch at +0.25 sees */
    TextureFilter getFilter() const { return (minFilter == LINEAR && magFilter == LINEAR) ? LINEAR : NEAREST; }
    ptr<DeviceContext> getContext() const { return context; }
    pl_pool *getPool() const { return pool; }
    size_t getTileBytes() const;
    /* the reference marks slots dirty for mipmap generation; nothing to do without mipmaps */
    void notifyChange(GPUSlot *s) { (void) s; }

    static bool parseInternalFormat(const std::string &name, TextureInternalFormat *f);
    static bool parseFilter(const std::string &name, TextureFilter *f);

protected:
    GPUTileStorage();
    void init(int tileSize, int nTiles, TextureInternalFormat internalf, TextureFilter min, TextureFilter mag,
              ptr<DeviceContext> context);

private:
    ptr<DeviceContext> context;
    pl_pool *pool;
    TextureInternalFormat internalf;
    TextureFilter minFilter, magFilter;
};

}  // namespace proland

#endif
