#include "proland/producer/GPUTileStorage.h"

#include <cassert>

namespace proland
{

GPUTileStorage::GPUSlot::GPUSlot(TileStorage *owner, int index, int l) : Slot(owner), l(l), index(index)
{
}

GPUTileStorage::GPUSlot::~GPUSlot()
{
}

int GPUTileStorage::GPUSlot::getWidth()
{
    return getOwner()->getTileSize();
}

int GPUTileStorage::GPUSlot::getHeight()
{
    return getOwner()->getTileSize();
}

void GPUTileStorage::GPUSlot::setSubImage(const void *pixels, size_t bytes)
{
    GPUTileStorage *s = static_cast<GPUTileStorage *>(getOwner());
    /* kernels queued earlier in this wave that read or write the slot go first (getImage flushes likewise) */
    s->getContext()->flush();
    DeviceContext::check(pl_pool_upload(s->getPool(), l, pixels, bytes));
}

void GPUTileStorage::GPUSlot::getImage(void *pixels, size_t bytes)
{
    GPUTileStorage *s = static_cast<GPUTileStorage *>(getOwner());
    s->getContext()->flush();
    DeviceContext::check(pl_pool_download(s->getPool(), l, pixels, bytes));
}

GPUTileStorage::GPUTileStorage(int tileSize, int nTiles, TextureInternalFormat internalf, TextureFilter min,
                               TextureFilter mag, ptr<DeviceContext> context) : TileStorage(), pool(NULL)
{
    init(tileSize, nTiles, internalf, min, mag, context);
}

GPUTileStorage::GPUTileStorage() : TileStorage(), pool(NULL), internalf(RGB32F), minFilter(NEAREST), magFilter(NEAREST)
{
}

void GPUTileStorage::init(int tileSize, int nTiles, TextureInternalFormat internalf, TextureFilter min,
                          TextureFilter mag, ptr<DeviceContext> context)
{
    TileStorage::init(tileSize, nTiles);
    this->context = context == NULL ? DeviceContext::get() : context;
    this->internalf = internalf;
    this->minFilter = min;
    this->magFilter = mag;
    int kind;
    switch (internalf) {
    case RGB32F:
    case RGBA32F: kind = PL_POOL_ELEV_F32x3; break;      /* (zf, zc, zm); the 4th channel of RGBA32F is always 0 */
    case RG8: kind = PL_POOL_NORM_UN8x2; break;
    case RGBA8: kind = PL_POOL_NORM_UN8x4; break;          /* normals (fine + coarse) or ortho colours */
    case RGB8: kind = PL_POOL_ORTHO_UN8x4; break;          /* ortho colours; stored as RGBA8, alpha unused */
    case R32F: kind = PL_POOL_RESID_F32; break;
    default: kind = PL_POOL_RESID_I16; break;
    }
    DeviceContext::check(pl_pool_create(this->context->handle(), kind, tileSize, nTiles, &pool));
    for (int i = 0; i < nTiles; ++i) {
        freeSlots.push_back(new GPUSlot(this, 0, i));
    }
}

GPUTileStorage::~GPUTileStorage()
{
    if (pool != NULL) {
        pl_pool_destroy(pool);
    }
}

const char *GPUTileStorage::getInternalFormatName() const
{
    switch (internalf) {
    case RGB32F: return "RGB32F";
    case RGBA32F: return "RGBA32F";
    case RG8: return "RG8";
    case RGBA8: return "RGBA8";
    case RGB8: return "RGB8";
    case R32F: return "R32F";
    default: return "R16I";
    }
}

int GPUTileStorage::getComponents() const
{
    switch (internalf) {
    case RGB32F: return 3;
    case RGBA32F: return 4;
    case RG8: return 2;
    case RGBA8: return 4;
    case RGB8: return 3;
    default: return 1;
    }
}

size_t GPUTileStorage::getTileBytes() const
{
    return pl_pool_tile_bytes(pool);
}

bool GPUTileStorage::parseInternalFormat(const std::string &name, TextureInternalFormat *f)
{
    static const struct { const char *n; TextureInternalFormat f; } table[] = {
        { "RGB32F", RGB32F }, { "RGBA32F", RGBA32F }, { "RG8", RG8 }, { "RGBA8", RGBA8 }, { "R32F", R32F }, { "R16I", R16I }, { "RGB8", RGB8 }
    };
    for (size_t i = 0; i < sizeof(table) / sizeof(table[0]); ++i) {
        if (name == table[i].n) {
            *f = table[i].f;
            return true;
        }
    }
    return false;
}

bool GPUTileStorage::parseFilter(const std::string &name, TextureFilter *f)
{
    if (name == "NEAREST") { *f = NEAREST; return true; }
    if (name == "LINEAR") { *f = LINEAR; return true; }
    /* mipmapped ortho storages (terrain3/helloworld.xml:43-45): the producers fetch level 0 at texel
     * centres, where every filter returns the texel */
    if (name == "LINEAR_MIPMAP_LINEAR" || name == "LINEAR_MIPMAP_NEAREST") { *f = LINEAR; return true; }
    if (name == "NEAREST_MIPMAP_NEAREST" || name == "NEAREST_MIPMAP_LINEAR") { *f = NEAREST; return true; }
    return false;
}

}  // namespace proland
