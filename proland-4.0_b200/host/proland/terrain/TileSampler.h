/*
 * TileSampler -- keeps, frame after frame, the tiles of one producer that the
 * quads of a terrain need: releases tiles of quads that disappeared, acquires
 * tiles of new quads, and returns the tasks that must run this frame.
 *
 * The update logic of the reference's TileSampler (core/sources/proland/terrain/
 * TileSampler.cpp:304-496): putTiles, getTiles, prefetch, the storeLeaf /
 * storeParent / async options and the Tree that mirrors the quadtree.  The GLSL
 * uniform side of the reference class (setTile, setTileMap: texture coordinates
 * for the renderer) is out of scope.
 */
#ifndef PROLAND_B200_TILE_SAMPLER_H
#define PROLAND_B200_TILE_SAMPLER_H

#include <string>

#include "proland/producer/TileProducer.h"
#include "proland/terrain/TerrainQuad.h"

namespace proland
{

PROLAND_API class TileSampler : public Object
{
public:
    TileSampler(const std::string &name, ptr<TileProducer> producer);
    virtual ~TileSampler();

    ptr<TileProducer> get() { return producer; }
    const std::string &getName() const { return name; }
    bool getStoreLeaf() const { return storeLeaf; }
    bool getStoreParent() const { return storeParent; }
    bool getAsync() const { return async; }
    void setStoreLeaf(bool v) { storeLeaf = v; }
    void setStoreParent(bool v) { storeParent = v; }
    /* async: tiles below the root are only taken when already in the cache, else prefetched */
    void setAsynchronous(bool v);

    /* one frame; the returned graph holds the tasks of the needed tiles that are not done.  frameNumber:
     * SceneManager::getFrameNumber() of the reference's update(scene, root) (used by TileSamplerZ) */
    virtual ptr<TaskGraph> update(ptr<TerrainQuad> root, unsigned int frameNumber = 0);
    /* tiles currently held (users taken by this sampler) */
    int getTileCount() const { return held; }
    /* drops every tile (call before the producer's cache goes away) */
    void release();

protected:
    struct Tree
    {
        bool newTree;
        bool needTile;
        Tree *parent;
        TileCache::Tile *t;
        Tree *children[4];
        explicit Tree(Tree *parent);
        virtual ~Tree();
    };

    std::string name;
    ptr<TileProducer> producer;
    Tree *root;
    bool storeLeaf;
    bool storeParent;
    bool async;
    int held;

    virtual bool needTile(ptr<TerrainQuad> q);
    virtual void recursiveDelete(Tree *t);
    void putTiles(Tree **t, ptr<TerrainQuad> q);
    virtual void getTiles(Tree *parent, Tree **t, ptr<TerrainQuad> q, ptr<TaskGraph> result);
    void prefetch(Tree *t, ptr<TerrainQuad> q, int &prefetchCount);
};

}  // namespace proland

#endif
