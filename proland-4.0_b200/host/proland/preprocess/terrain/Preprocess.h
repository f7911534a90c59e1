/*
 * Preprocess -- the residual-file builder: a height map -> DEM.dat (flat) or DEM1..6.dat (cube faces) in the format
 * ResidualProducer reads.
 *
 * Host mirror of terrain/sources/proland/preprocess/terrain/Preprocess.h:58-215 (InputMap, preprocessDem,
 * preprocessSphericalDem) with the same signatures.  Different underneath: the reference samples the source map per
 * pixel through an LRU tile cache, writes every mipmap level and every approximation tile to temporary TIFF / raw
 * files and reads them back (HeightMipmap.cpp:132-324); here the map is uploaded once, the base level of every face
 * stays resident in HBM (pl_height_cube_*), and each level is three batched device passes -- height tiles gathered
 * across the cube edges (pl_height_tiles), residual + approximation (pl_residual_encode_batch) -- with the
 * approximations of the previous level kept in a device pool.  tmpFolder is accepted and unused.
 */
#ifndef PROLAND_B200_PREPROCESS_H
#define PROLAND_B200_PREPROCESS_H

#include <string>

#include "ork/ork_lite.h"

using namespace std;
using namespace ork;

namespace proland
{

/* Preprocess.h:58-154: an abstract raster, read by pixel or by tile */
PROLAND_API class InputMap
{
public:
    int width;
    int height;
    int channels;
    int tileSize;

    InputMap(int width, int height, int channels, int tileSize, int cache = 20);
    virtual ~InputMap();

    virtual vec4f getValue(int x, int y) = 0;
    /* tileSize x tileSize x channels values of the tile whose lower-left pixel is (x, y); new[]-allocated */
    virtual float *getValues(int x, int y);
    /* the pixel (x, y), coordinates clamped to the map (Preprocess.cpp:129-152) */
    vec4f get(int x, int y);
};

/* An InputMap over a float array the caller keeps alive (row 0 first, one channel): what the reference's preprocess examples
 * subclass InputMap for.  Not in the reference; the tile size is the largest power of two <= 256 dividing both extents. */
class ArrayInputMap : public InputMap
{
public:
    ArrayInputMap(const float *data, int width, int height);
    virtual vec4f getValue(int x, int y);
    virtual float *getValues(int x, int y);

private:
    const float *data;
    static int tileFor(int width, int height);
};

/* Preprocess.cpp:512-533 */
PROLAND_API void preprocessDem(InputMap *src, int dstMinTileSize, int dstTileSize, int dstMaxLevel,
    const string &dstFolder, const string &tmpFolder, float residualScale);

/* Preprocess.cpp:535-585 */
PROLAND_API void preprocessSphericalDem(InputMap *src, int dstMinTileSize, int dstTileSize, int dstMaxLevel,
    const string &dstFolder, const string &tmpFolder, float residualScale);

}

#endif
