/* Shim standing in for libtiff's tiffio.h (libtiff 3.x is a binary-only dependency of the reference: bin/libtiff3.dll;
 * no headers or library in this image) so that the reference's preprocess/terrain sources compile UNCHANGED.
 * Test infrastructure: used only to run the reference's own residual builder (HeightMipmap) as a checker.
 *
 * What it provides is the handful of calls those sources make, over files kept IN MEMORY:
 *   TIFFOpen(name, "wb"/"rb") + TIFFSetField + TIFFWriteEncodedStrip + TIFFWriteDirectory + TIFFSetDirectory +
 *   TIFFReadEncodedStrip + TIFFClose        the builder's temporary mipmap / residual tile files: a named list of
 *                                           directories, each one strip of raw bytes (lossless, as DEFLATE is)
 *   TIFFClientOpen(... mfs procs ...)       the tile blobs of the final container: on TIFFClose a little-endian TIFF
 *                                           (header, one zlib strip at byte 8, IFD with the tags SURVEY 8c lists for
 *                                           terrain4/DEM.dat) goes through the client's write procedure
 */
#ifndef ORC_SHIM_TIFFIO_H
#define ORC_SHIM_TIFFIO_H
#include <cstdarg>
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>

typedef int tsize_t;
typedef void *tdata_t;
typedef void *thandle_t;
typedef int toff_t;
typedef unsigned int ttag_t;
typedef unsigned int tstrip_t;
typedef unsigned short tdir_t;
typedef tsize_t (*TIFFReadWriteProc)(thandle_t, tdata_t, tsize_t);
typedef toff_t (*TIFFSeekProc)(thandle_t, toff_t, int);
typedef int (*TIFFCloseProc)(thandle_t);
typedef toff_t (*TIFFSizeProc)(thandle_t);
typedef int (*TIFFMapFileProc)(thandle_t, tdata_t *, toff_t *);
typedef void (*TIFFUnmapFileProc)(thandle_t, tdata_t, toff_t);

#define TIFFTAG_IMAGEWIDTH 256
#define TIFFTAG_IMAGELENGTH 257
#define TIFFTAG_BITSPERSAMPLE 258
#define TIFFTAG_COMPRESSION 259
#define TIFFTAG_PHOTOMETRIC 262
#define TIFFTAG_STRIPOFFSETS 273
#define TIFFTAG_ORIENTATION 274
#define TIFFTAG_SAMPLESPERPIXEL 277
#define TIFFTAG_ROWSPERSTRIP 278
#define TIFFTAG_STRIPBYTECOUNTS 279
#define TIFFTAG_PLANARCONFIG 284
#define TIFFTAG_JPEGQUALITY 65537
#define TIFFTAG_JPEGCOLORMODE 65538
#define JPEGCOLORMODE_RGB 1
#define COMPRESSION_NONE 1
#define COMPRESSION_JPEG 7
#define COMPRESSION_DEFLATE 32946
#define ORIENTATION_TOPLEFT 1
#define ORIENTATION_BOTLEFT 4
#define PLANARCONFIG_CONTIG 1
#define PHOTOMETRIC_MINISBLACK 1
#define PHOTOMETRIC_RGB 2
#define PHOTOMETRIC_YCBCR 6

struct orc_tiff_dir {
    std::map<unsigned, int> fields;
    std::vector<unsigned char> strip;
};
struct TIFF {
    std::string name;
    bool writing, client;
    std::vector<orc_tiff_dir> *dirs;    /* named file: the store's entry */
    orc_tiff_dir cur;                   /* directory being written */
    size_t sel;                         /* directory selected for reading */
    thandle_t handle;
    TIFFReadWriteProc wr;
    TIFFCloseProc cl;
};

extern "C" {
TIFF *TIFFOpen(const char *name, const char *mode);
TIFF *TIFFClientOpen(const char *name, const char *mode, thandle_t h, TIFFReadWriteProc rd, TIFFReadWriteProc wr, TIFFSeekProc sk,
                     TIFFCloseProc cl, TIFFSizeProc sz, TIFFMapFileProc mp, TIFFUnmapFileProc um);
int TIFFSetField(TIFF *t, ttag_t tag, ...);
int TIFFGetField(TIFF *t, ttag_t tag, ...);
tsize_t TIFFWriteEncodedStrip(TIFF *t, tstrip_t strip, tdata_t data, tsize_t size);
tsize_t TIFFReadEncodedStrip(TIFF *t, tstrip_t strip, tdata_t data, tsize_t size);
int TIFFWriteDirectory(TIFF *t);
int TIFFSetDirectory(TIFF *t, tdir_t n);
void TIFFClose(TIFF *t);
/* the shim's own: forget every named in-memory file */
void orc_tiff_store_clear(void);
}
#endif
