/* extern "C" face of the reference's own residual builder: preprocess/terrain/{Preprocess, HeightMipmap, AbstractTileCache,
 * ColorMipmap, Util}.cpp and core/.../util/mfs.cpp, compiled UNCHANGED into oracle/_ref/libref_hm.so by oracle/Makefile over
 * the shims in ref_shim/hm (Ork's Object.h / vec3.h / vec4.h, and an in-memory tiffio.h: libtiff is a binary-only
 * dependency of the reference).  Test infrastructure: used only to pin oracle/orc_preprocess.c and to check the device
 * builder (pl_height_* + pl_residual_encode_batch + pl_residual_write_file) against files the reference's code wrote.
 *
 * This file holds (1) the implementation of the tiffio.h shim and (2) two entry points that run
 * proland::preprocessSphericalDem / proland::preprocessDem on a float map given by the caller. */
#include <fcntl.h>
#include <zlib.h>

#include "tiffio.h"
#include "proland/preprocess/terrain/Preprocess.h"

/* ------------------------------------------------------------------------------------------------- tiffio.h shim */
static std::map<std::string, std::vector<orc_tiff_dir> > g_store;

extern "C" void orc_tiff_store_clear(void) { g_store.clear(); }

extern "C" TIFF *TIFFOpen(const char *name, const char *mode)
{
    const bool writing = mode[0] == 'w';
    if (!writing && g_store.find(name) == g_store.end()) return NULL;
    TIFF *t = new TIFF();
    t->name = name;
    t->writing = writing;
    t->client = false;
    if (writing) g_store[name].clear();
    t->dirs = &g_store[name];
    t->sel = 0;
    return t;
}

extern "C" TIFF *TIFFClientOpen(const char *name, const char *mode, thandle_t h, TIFFReadWriteProc, TIFFReadWriteProc wr, TIFFSeekProc,
                                TIFFCloseProc cl, TIFFSizeProc, TIFFMapFileProc, TIFFUnmapFileProc)
{
    if (mode[0] != 'w') return NULL;      /* the builder only writes through client procedures */
    TIFF *t = new TIFF();
    t->name = name;
    t->writing = true;
    t->client = true;
    t->dirs = NULL;
    t->sel = 0;
    t->handle = h;
    t->wr = wr;
    t->cl = cl;
    return t;
}

extern "C" int TIFFSetField(TIFF *t, ttag_t tag, ...)
{
    va_list ap;
    va_start(ap, tag);
    t->cur.fields[tag] = va_arg(ap, int);
    va_end(ap);
    return 1;
}

extern "C" int TIFFGetField(TIFF *t, ttag_t tag, ...)
{
    va_list ap;
    va_start(ap, tag);
    int *out = va_arg(ap, int *);
    va_end(ap);
    if (!t->dirs || t->sel >= t->dirs->size()) return 0;
    const std::map<unsigned, int> &f = (*t->dirs)[t->sel].fields;
    std::map<unsigned, int>::const_iterator i = f.find(tag);
    if (i == f.end()) return 0;
    *out = i->second;
    return 1;
}

extern "C" tsize_t TIFFWriteEncodedStrip(TIFF *t, tstrip_t, tdata_t data, tsize_t size)
{
    t->cur.strip.assign((const unsigned char *) data, (const unsigned char *) data + size);
    return size;
}

extern "C" int TIFFWriteDirectory(TIFF *t)
{
    if (t->dirs) t->dirs->push_back(t->cur);
    t->cur = orc_tiff_dir();
    return 1;
}

extern "C" int TIFFSetDirectory(TIFF *t, tdir_t n)
{
    if (!t->dirs || n >= t->dirs->size()) return 0;
    t->sel = n;
    return 1;
}

extern "C" tsize_t TIFFReadEncodedStrip(TIFF *t, tstrip_t, tdata_t data, tsize_t size)
{
    if (!t->dirs || t->sel >= t->dirs->size()) return -1;
    const std::vector<unsigned char> &s = (*t->dirs)[t->sel].strip;
    const size_t n = size < 0 ? s.size() : std::min((size_t) size, s.size());
    memcpy(data, s.data(), n);
    return (tsize_t) n;
}

static void put16(std::vector<unsigned char> &b, unsigned v) { b.push_back(v & 255); b.push_back((v >> 8) & 255); }
static void put32(std::vector<unsigned char> &b, unsigned v) { put16(b, v & 0xFFFF); put16(b, v >> 16); }
static void put_tag(std::vector<unsigned char> &b, unsigned tag, unsigned type, unsigned count, unsigned value)
{
    put16(b, tag); put16(b, type); put32(b, count);
    if (type == 3 && count == 1) { put16(b, value); put16(b, 0); } else put32(b, value);
}

/* a client file: header, one zlib strip at byte 8, then the IFD (the layout of the blobs in terrain4/DEM.dat) */
static void write_client_tiff(TIFF *t)
{
    const orc_tiff_dir &d = t->cur;
    uLongf zlen = compressBound(d.strip.size());
    std::vector<unsigned char> z(zlen);
    compress2(z.data(), &zlen, d.strip.data(), d.strip.size(), Z_DEFAULT_COMPRESSION);
    std::vector<unsigned char> b;
    b.push_back('I'); b.push_back('I'); put16(b, 42);
    unsigned ifd = 8 + (unsigned) zlen;
    ifd += ifd & 1;
    put32(b, ifd);
    b.insert(b.end(), z.begin(), z.begin() + zlen);
    if (b.size() & 1) b.push_back(0);
    std::map<unsigned, int> f = d.fields;
    const unsigned spp = f.count(TIFFTAG_SAMPLESPERPIXEL) ? f[TIFFTAG_SAMPLESPERPIXEL] : 1;
    const unsigned bps = f.count(TIFFTAG_BITSPERSAMPLE) ? f[TIFFTAG_BITSPERSAMPLE] : 8;
    put16(b, 10);
    put_tag(b, TIFFTAG_IMAGEWIDTH, 3, 1, f[TIFFTAG_IMAGEWIDTH]);
    put_tag(b, TIFFTAG_IMAGELENGTH, 3, 1, f[TIFFTAG_IMAGELENGTH]);
    put16(b, TIFFTAG_BITSPERSAMPLE); put16(b, 3); put32(b, spp); put16(b, bps); put16(b, spp > 1 ? bps : 0);
    put_tag(b, TIFFTAG_COMPRESSION, 3, 1, f[TIFFTAG_COMPRESSION]);
    put_tag(b, TIFFTAG_PHOTOMETRIC, 3, 1, f[TIFFTAG_PHOTOMETRIC]);
    put_tag(b, TIFFTAG_STRIPOFFSETS, 4, 1, 8);
    put_tag(b, TIFFTAG_ORIENTATION, 3, 1, f[TIFFTAG_ORIENTATION]);
    put_tag(b, TIFFTAG_SAMPLESPERPIXEL, 3, 1, spp);
    put_tag(b, TIFFTAG_STRIPBYTECOUNTS, 4, 1, (unsigned) zlen);
    put_tag(b, TIFFTAG_PLANARCONFIG, 3, 1, f[TIFFTAG_PLANARCONFIG]);
    put32(b, 0);
    t->wr(t->handle, b.data(), (tsize_t) b.size());
}

extern "C" void TIFFClose(TIFF *t)
{
    if (!t) return;
    if (t->client) {
        write_client_tiff(t);
        if (t->cl) t->cl(t->handle);
    }
    delete t;
}

/* ------------------------------------------------------------------------------------------------- entry points */
namespace {

class ArrayMap : public proland::InputMap
{
public:
    const float *data;
    ArrayMap(const float *data, int w, int h, int tile) : proland::InputMap(w, h, 1, tile), data(data) {}
    virtual vec4f getValue(int x, int y) { return vec4f(data[(size_t) y * width + x], 0, 0, 0); }
};

int pick_tile(int w, int h)
{
    for (int t = 256; t > 1; t /= 2)
        if (w % t == 0 && h % t == 0) return t;
    return 1;
}

/* the reference's progress lines go to stdout: silence them for the duration of a call */
struct Quiet {
    int saved;
    Quiet() { fflush(stdout); saved = dup(1); int n = open("/dev/null", O_WRONLY); dup2(n, 1); close(n); }
    ~Quiet() { fflush(stdout); dup2(saved, 1); close(saved); }
};

}  // namespace

extern "C" {

/* proland::preprocessSphericalDem (Preprocess.cpp:535-585): src = an equirectangular height map (sw x sh floats);
 * writes dst/DEM1.dat .. DEM6.dat; tmp: a scratch directory prefix (the builder keeps its .raw approximation tiles there) */
int ref_preprocess_spherical_dem(const float *src, int sw, int sh, int minTileSize, int tileSize, int maxLevel, const char *dst,
                                 const char *tmp, float scale)
{
    Quiet q;
    orc_tiff_store_clear();
    ArrayMap m(src, sw, sh, pick_tile(sw, sh));
    try {
        proland::preprocessSphericalDem(&m, minTileSize, tileSize, maxLevel, dst, tmp, scale);
    } catch (...) {
        orc_tiff_store_clear();
        return -1;
    }
    orc_tiff_store_clear();
    return 0;
}

/* proland::preprocessDem (Preprocess.cpp:512-533): a flat DEM -> dst/DEM.dat */
int ref_preprocess_dem(const float *src, int sw, int sh, int minTileSize, int tileSize, int maxLevel, const char *dst, const char *tmp,
                       float scale)
{
    Quiet q;
    orc_tiff_store_clear();
    ArrayMap m(src, sw, sh, pick_tile(sw, sh));
    try {
        proland::preprocessDem(&m, minTileSize, tileSize, maxLevel, dst, tmp, scale);
    } catch (...) {
        orc_tiff_store_clear();
        return -1;
    }
    orc_tiff_store_clear();
    return 0;
}

}
