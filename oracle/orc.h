/*
 * orc.h -- CPU ORACLE for the Proland 4.0 terrain tile-production hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * algorithm (GLSL shaders + the C++ host code that feeds them).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  The product path (proland-4.0_b200/) never links, imports or
 * calls anything in this directory.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * the reference checkout root).
 *
 * Pinning status (see DESIGN.md "Oracle"):
 *   - LCG / frandom / cnoise / createDemNoise : pinned bit-exactly against the
 *     reference's own noise.cpp compiled unchanged (oracle/_ref) and against
 *     the KATs recorded in SURVEY.md 8c.
 *   - residual container decode : pinned against the reference's fixture
 *     src/terrain/examples/terrain4/DEM.dat (sha1 KATs) via stock zlib.
 *   - 4x4 upsample filter : pinned against the reference's own CPU restatement
 *     (CPUElevationProducer.cpp:194-245, re-stated in orc_cpu_elevation_tile).
 *   - the float arithmetic of the three shaders (orc_upsample_tile, orc_normal_tile, orc_ortho_tile): pinned
 *     on the reference's own GLSL text, compiled unchanged as C++ behind oracle/ref_shim/glsl_shim.h
 *     (oracle/_ref/libref_glsl.so, oracle/ref_glsl_wrap.cpp).  The non-contracted build of this restatement
 *     (liborc_strict.so) equals it BIT FOR BIT on every case of tests/glsl_cases.py (all shader variants, the
 *     deep chains of BASELINE configs 1-4, four normal formats, ortho hsv / plain / residuals); the hashes are
 *     committed (tests/golden/glsl.json).  The canonical build (liborc.so, a*b+c fused -- the order the CUDA
 *     kernels share) differs from the text only by that contraction, which GLSL 3.30 leaves to the
 *     implementation: measured in tests/test_glsl_pin.py, <= 1e-5 of the height range, <= 1 unorm8 step.
 *   - what GL itself does around the shaders stays taken from the OpenGL 3.3 spec (no GL here): fp16 rounding
 *     of the R16F upload, unorm8 rounding, clamp-to-edge, LINEAR weights with 8 subtexel bits, and Ork's
 *     vec3d::normalize.
 */
#ifndef ORC_H
#define ORC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ noise */

/* core/sources/proland/math/noise.h:50-54 */
long orc_lrandom(long *seed);
/* core/sources/proland/math/noise.h:63-67 */
float orc_frandom(long *seed);
/* core/sources/proland/math/noise.cpp:117-165 (2D classic Perlin, period 0) */
float orc_cnoise2(float x, float y);
/* copy of the 2D gradient table + permutation (noise.cpp:68-101), for upload
 * to the device: p[514] ints, g2[514*2] floats */
void orc_cnoise_tables(int *p, float *g2);

/* terrain/sources/proland/dem/ElevationProducer.cpp:50-128; out = 6*W*W fp32 */
void orc_dem_noise(int W, float *out);
/* fp32 -> fp16 -> fp32, round-to-nearest-even: the R16F upload of
 * ElevationProducer.cpp:129 (rounding mode per OpenGL 3.3 spec 2.1.2) */
float orc_round_half(float v);
uint16_t orc_float_to_half_bits(float v);
/* orc_dem_noise followed by orc_round_half on every texel */
void orc_dem_noise_r16f(int W, float *out);

/* ElevationProducer.cpp:345-373 : noise layer / rotation selection */
void orc_noise_select(int level, int tx, int ty, int face, int *noiseR, int *noiseL);

/* -------------------------------------------------------------- elevation */

enum { ORC_NOISE_PLAIN = 0,   /* variants A, C: zf += |rs| * n              */
       ORC_NOISE_SLOPE = 1 }; /* variants B, D: slope/curvature modulated    */

typedef struct {
    int W;              /* tile width incl. borders (tileWSDF.x)              */
    int level;
    float pixel_size;   /* tileWSDF.y                                          */
    int grid;           /* tileWSDF.z = (W-5)/gridMeshSize (integer division)  */
    int flip;           /* tileWSDF.w > 0 and the shader variant honours it    */
    int dx, dy;         /* coarseLevelOSL.xy in texels: (t%2)*(W-5)/2          */
    int has_resid;      /* residualOSH.w == 1                                  */
    int rx, ry;         /* residual window origin ElevationProducer.cpp:324    */
    int resid_stride;   /* residual tile width (197)                           */
    int noiseR, noiseL; /* noiseUVLH.x, .z                                     */
    float rs;           /* noiseUVLH.w                                         */
    int noise_mode;     /* ORC_NOISE_*                                         */
    int no_clamp;       /* #define NO_CLAMP                                    */
} orc_elev_params;

/* Fill the per-tile uniform block exactly as ElevationProducer::doCreateTile
 * does (ElevationProducer.cpp:293-376). resid_W = 0 when no residual producer */
void orc_elev_uniforms(int W, int gridMeshSize, float rootQuadSize, int flip,
                       const float *noiseAmp, int nAmp, int face,
                       int level, int tx, int ty,
                       int has_resid, int resid_W,
                       int noise_mode, int no_clamp, orc_elev_params *p);

/* src/demo/shaders/elevation/upsampleShader.glsl:140-203 (variant D) and the
 * example variants A/B/C.  parent: W*W*3 interleaved (zf,zc,zm) or
 * NULL at level 0.  resid: residual tile (stride p->resid_stride) or NULL.
 * noise: 6*W*W, already fp16-rounded.  out: W*W*3 interleaved. */
void orc_upsample_tile(const orc_elev_params *p, const float *parent,
                       const float *resid, const float *noise, float *out);

/* terrain/sources/proland/dem/CPUElevationProducer.cpp:194-245 (one channel) */
void orc_cpu_elevation_tile(int W, int level, int tx, int ty, const float *parent,
                            const float *resid, int resid_W, int rx, int ry, float *out);

/* core/sources/proland/terrain/TileSamplerZ.cpp:43-133 : min/max of zm over
 * texels [2, W-3]^2 */
void orc_tile_minmax(int W, const float *elev, float *zmin, float *zmax);

/* ---------------------------------------------------------------- normals */

enum { ORC_FILTER_NEAREST = 0, ORC_FILTER_LINEAR = 1 };

typedef struct {
    int W;              /* normal tile width (tileSDF.x) = 97                  */
    int grid;           /* tileSDF.y = (W-1)/gridMeshSize                      */
    int format;         /* tileSDF.z : 0 RGBA signed, 1 RGBA unsigned,
                                       2 RG signed,   3 RG unsigned            */
    int elev_W;         /* elevation tile width (101)                          */
    int elev_border;    /* 2                                                   */
    int elev_filter;    /* min/mag filter of the elevation storage             */
    int has_parent;     /* normalOSL.x != -1 (level>0 and 4 components)        */
    int ptx, pty;       /* tx%2, ty%2                                          */
    int parent_filter;  /* min/mag filter of the normal storage                */
    float deform[4];    /* x0, y0, quad size, R (0 = flat)                     */
    float corners[16];  /* patchCorners, row-major maths matrix                */
    float verticals[16];
    float norms[4];
    float w2t[9];       /* worldToTangentFrame row-major                       */
    float p2t[9];       /* parentToTangentFrame row-major                      */
} orc_norm_params;

/* NormalProducer.cpp:175-283 (double precision host maths -> fp32 uniforms) */
void orc_normal_uniforms(int W, int gridMeshSize, int components, int signed_comp,
                         int elev_W, int elev_border, int elev_filter, int parent_filter,
                         double rootQuadSize, int sphere,
                         int level, int tx, int ty, orc_norm_params *p);

/* src/demo/shaders/elevation/normalShader.glsl:60-125.
 * elev: elev_W*elev_W*3 interleaved.  parent: parent normal tile as the
 * sampler returns it (W*W*4 floats, texel values in [0,1] for unorm storage)
 * or NULL.  out: W*W*4 floats = the fragment's `data` before framebuffer
 * conversion. */
void orc_normal_tile(const orc_norm_params *p, const float *elev,
                     const float *parent, float *out);

/* OpenGL 3.3 spec 2.1.5: float -> unorm8 = round(clamp(f,0,1)*255) */
uint8_t orc_unorm8(float f);
/* data -> RG8 (2 bytes/px) or RGBA8 (4 bytes/px) */
void orc_pack_unorm8(int W, int channels, const float *data, uint8_t *out);

/* -------------------------------------------------------------- residuals */

typedef struct orc_resid_file {
    int minLevel, maxLevel, tileSize, rootLevel, rootTx, rootTy;
    float scale;          /* file scale * zscale                               */
    int deltaLevel;
    int ntiles;
    uint32_t header;      /* byte offset of the first blob                     */
    const uint32_t *offsets;
    const uint8_t *data;  /* whole file, caller-owned                          */
    size_t size;
    int nchildren;
    struct orc_resid_file **children; /* nested <residualProducer> elements   */
} orc_resid_file;

/* ResidualProducer.cpp:70-129. returns 0 on success */
int orc_resid_open(const uint8_t *data, size_t size, int deltaLevel, float zscale,
                   orc_resid_file *f);
/* ResidualProducer.cpp:161-175 */
int orc_resid_has_tile(const orc_resid_file *f, int level, int tx, int ty);
/* ResidualProducer.cpp:253-266 */
int orc_resid_tile_size(const orc_resid_file *f, int l);
int orc_resid_tile_id(const orc_resid_file *f, int l, int tx, int ty);
/* blob location of a stored tile id: returns pointer + size */
const uint8_t *orc_resid_blob(const orc_resid_file *f, int tileid, uint32_t *size);
/* TIFF (1 strip, DEFLATE) -> raw bytes; what TIFFReadEncodedStrip returns
 * (ResidualProducer.cpp:312-319).  Returns number of bytes or <0 on error */
long orc_tiff_inflate(const uint8_t *blob, uint32_t size, uint8_t *raw, size_t cap,
                      int *width, int *height);
/* ResidualProducer.cpp:268-340 (stored-level coordinates) */
int orc_resid_read_tile(const orc_resid_file *f, int l, int tx, int ty,
                        const float *tile, float *result);
/* ResidualProducer.cpp:342-384 */
void orc_resid_upsample(const orc_resid_file *f, int l, int tx, int ty,
                        const float *parentTile, float *result);
/* ResidualProducer.cpp:177-232 : full doCreateTile (absolute coordinates).
 * out: (tileSize+5)^2 floats */
int orc_resid_create_tile(const orc_resid_file *f, int level, int tx, int ty, float *out);

/* ----------------------------------------------------------------- driver */

typedef struct {
    int W;                 /* 101 */
    int gridMeshSize;      /* 24  */
    float rootQuadSize;
    int face;
    int flip;
    int noise_mode;
    int no_clamp;
    int nAmp;
    float noiseAmp[32];
    int sphere;            /* NormalProducer deform="sphere" */
    int elev_filter;       /* elevation storage min/mag filter */
    const orc_resid_file *resid; /* or NULL */
} orc_scene;

/* Produce one elevation tile (W*W*3) + its RG8 normal tile ((W-4)^2*2) the way
 * the reference's two producers do.  parent = parent elevation tile or NULL.
 * resid_tile = the residual tile covering (level,tx,ty) or NULL.
 * noise = orc_dem_noise_r16f(W). */
void orc_produce_pair(const orc_scene *s, const float *noise, int level, int tx, int ty,
                      const float *parent, const float *resid_tile,
                      float *elev_out, uint8_t *norm_out);

/* Full quadtree, levels 0..maxLevel of the subtree under (rootLevel 0), BFS by
 * level, OpenMP over the tiles of a level (nthreads<=0: all).  Keeps only two
 * levels alive.  Returns the number of pairs produced; fills a checksum
 * (sum over tiles of zmin+zmax, double) and global zmin/zmax. */
long orc_produce_quadtree(const orc_scene *s, int maxLevel, int nthreads,
                          double *checksum, float *zmin, float *zmax);

/* -------------------------------------------------------------------- ortho
 * OrthoProducer, orc_ortho.c */
typedef struct orc_ortho_params {
    int tileWidth;           /* tileWidth uniform (storage tile size, e.g. 196)          */
    int level;
    int dx, dy;              /* coarseLevelOSL.xy in texels, -1 at level 0                */
    int hasResidual;         /* residualOSH.x != -1                                       */
    float residualScale;     /* residualOSH.w (scale, or -1 without residual)             */
    int noiseR, noiseL;      /* noiseUVLH.x, .z                                           */
    int hsv;                 /* noiseUVLH.w                                               */
    float noiseColor[4];     /* noiseColor uniform (OrthoProducer.cpp:357-361)            */
    float rootNoiseColor[4];
} orc_ortho_params;
/* OrthoProducer.cpp:48-118; out = 6*W*W*4 bytes */
void orc_ortho_noise(int W, uint8_t *out);
/* OrthoProducer.cpp:286-372 */
void orc_ortho_uniforms(int W, int face, int level, int tx, int ty, const float *noiseAmp, int nAmp,
                        const float noiseColor[4], const float rootNoiseColor[4], int hsv, float scale,
                        int hasResidual, orc_ortho_params *p);
/* upsampleOrthoShader.glsl:123-158 */
void orc_ortho_tile(const orc_ortho_params *p, const uint8_t *parent, const uint8_t *residual, int channels,
                    const uint8_t *noise, uint8_t *out);
/* OrthoCPUProducer.cpp:68-118,160-246: one tile of an ortho residual file -> W*W*channels bytes; returns W or < 0 */
int orc_ortho_cpu_read(const uint8_t *file, size_t size, int level, int tx, int ty, uint8_t *out, int *channels);
long orc_ortho_quadtree(int W, int face, int maxLevel, const float *noiseAmp, int nAmp, const float noiseColor[4],
                        const float rootNoiseColor[4], int hsv, float scale, const uint8_t *noise, uint8_t *out);

/* the height pyramid of a cube: HeightMipmap::getTileHeight / getTile with setCube's stitching, the six cube
 * projections and SphericalHeightFunction (orc_preprocess.c) */
float orc_hm_height(const short *const *faces, int nfaces, int B, int maxLevel, int level, int face, int x, int y);
void orc_hm_get_tile(const short *const *faces, int nfaces, int B, int maxLevel, int topLevelSize, int tileSize, float scale,
                     int level, int face, int tx, int ty, float *tile);
void orc_cube_projection(int face, int x, int y, int w, double *sx, double *sy, double *sz);
short orc_spherical_base_sample(const float *src, int sw, int sh, int face, int x, int y, int B);
void orc_spherical_base_grid(const float *src, int sw, int sh, int face, int B, short *out);
void orc_plane_base_grid(const float *src, int sw, int sh, int B, short *out);

/* ---------------------------------------------------------------- preprocess
 * the residual-pyramid builder, one tile of one level (HeightMipmap.cpp:449-559), orc_preprocess.c */
void orc_hm_encode_tile(const float *parentTile, const float *tile, int n, int tileSize, int tx, int ty,
                        short *resid, float *approx, float *maxR, float *maxErr);

#ifdef __cplusplus
}
#endif

#endif
