/*
 * TileSamplerZ -- a TileSampler for elevation tiles that also reads back, a few frames late, the z range of the
 * tiles it holds into TerrainQuad::zmin / zmax and the ground height under the camera into
 * TerrainNode::groundHeightAtCamera, which the split rule of the quadtree consumes (TerrainQuad.cpp:102-105).
 *
 * The reference (core/sources/proland/terrain/TileSamplerZ.cpp:43-133, 253-386) computes the min / max of a tile with
 * a mipmapping shader at read-back time and fetches 16 results per frame through a ReadbackManager(1, 3, ..) -- one
 * read-back per frame, three in flight.  Here the elevation kernel has already left (zmin, zmax) of every tile it
 * produced in the pool (the fused epilogue, same texel range [2, W-3]^2), so a frame only GATHERS: the slots it picked
 * and the texel under the camera go through pl_elev_zreadback_begin, the result is collected two frames later.
 * Same state machine otherwise: needReadback ordered by (level, address), at most 16 tiles per frame, the camera
 * slot first, one read-back per frame and storage however many samplers share it, camera quads always need a tile.
 */
#ifndef PROLAND_B200_TILE_SAMPLER_Z_H
#define PROLAND_B200_TILE_SAMPLER_Z_H

#include <deque>
#include <map>
#include <set>
#include <vector>

#include "proland/producer/GPUTileStorage.h"
#include "proland/terrain/TileSampler.h"

namespace proland
{

PROLAND_API class TileSamplerZ : public TileSampler
{
public:
    /* tiles read back per frame (MAX_MIPMAP_PER_FRAME) and frames a result stays in flight before it is applied */
    enum { MAX_TILES_PER_FRAME = 16, READBACK_DELAY = 2 };

    TileSamplerZ(const std::string &name, ptr<TileProducer> producer);
    virtual ~TileSamplerZ();

    virtual ptr<TaskGraph> update(ptr<TerrainQuad> root, unsigned int frameNumber = 0);
    /* read-backs issued / applied so far, tiles waiting for one (tests) */
    void getCounts(unsigned long long out[3]) const;

protected:
    struct TreeZ : public Tree
    {
        ptr<TerrainQuad> q;
        bool readback;               /* a read-back of this tile has been requested */
        unsigned int readbackDate;   /* completion date of the tile content that request was for */
        TreeZ(Tree *parent, ptr<TerrainQuad> q);
    };
    struct TreeZSort
    {
        bool operator()(const TreeZ *x, const TreeZ *y) const;
    };

    /* shared by the samplers of one storage (stateFactory of the reference) */
    struct State
    {
        GPUTileStorage *storage;
        int users;
        std::set<TreeZ *, TreeZSort> needReadback;
        GPUTileStorage::GPUSlot *cameraSlot;
        int cameraX, cameraY;
        unsigned int lastFrame;
        struct Pending { int ticket; unsigned int frame; bool camera; std::vector<ptr<TerrainQuad> > targets; };
        std::deque<Pending> pending;
        unsigned long long issued, applied;
    };

    virtual bool needTile(ptr<TerrainQuad> q);
    virtual void recursiveDelete(Tree *t);
    virtual void getTiles(Tree *parent, Tree **t, ptr<TerrainQuad> q, ptr<TaskGraph> result);

private:
    State *state;
    TreeZ *cameraQuad;
    float cameraQuadX, cameraQuadY;
    double oldCamX, oldCamY, oldCamZ;

    void collect(unsigned int frameNumber, bool all);
    static std::map<GPUTileStorage *, State *> states;
};

}  // namespace proland

#endif
