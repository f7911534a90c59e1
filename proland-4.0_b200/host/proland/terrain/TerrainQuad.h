/*
 * TerrainNode / TerrainQuad -- the view-dependent quadtree that decides which tiles
 * a frame needs: the caller side of the tile-production path.
 *
 * Same subdivision rule, child order and update order as the reference
 * (core/sources/proland/terrain/TerrainQuad.cpp:81-173, TerrainNode.cpp:116-145):
 * a quad is split iff dist < l * splitDist and level < maxLevel, where dist is
 * TerrainNode::getCameraDist of the quad's box at ground height; children are
 * (2tx,2ty) (2tx+1,2ty) (2tx,2ty+1) (2tx+1,2ty+1) and are updated nearest-first
 * by camera quadrant.  Frustum / horizon culling (SceneManager, Deformation) is
 * renderer work and not carried over: every quad is visible, which only makes the
 * tile set a superset of the reference's.
 */
#ifndef PROLAND_B200_TERRAIN_QUAD_H
#define PROLAND_B200_TERRAIN_QUAD_H

#include "ork/ork_lite.h"

using namespace ork;

namespace proland
{

class TerrainNode;

PROLAND_API class TerrainQuad : public Object
{
public:
    const TerrainQuad *parent;
    const int level;
    const int tx;
    const int ty;
    const double ox;      /* lower-left corner, local space */
    const double oy;
    const double l;       /* side length */
    float zmin;
    float zmax;
    ptr<TerrainQuad> children[4];

    TerrainQuad(TerrainNode *owner, const TerrainQuad *parent, int tx, int ty, double ox, double oy, double l,
                float zmin, float zmax);
    virtual ~TerrainQuad();

    TerrainNode *getOwner();
    bool isLeaf() const;
    int getSize() const;      /* quads in this subtree */
    int getDepth() const;     /* deepest level below */
    void update();

private:
    TerrainNode *owner;
    void subdivide();
    friend class TerrainNode;
};

PROLAND_API class TerrainNode : public Object
{
public:
    ptr<TerrainQuad> root;
    int maxLevel;
    /* height of the ground under the camera (TerrainNode::groundHeightAtCamera, fed by TileSamplerZ) */
    static float groundHeightAtCamera;
    /* the value groundHeightAtCamera will have at the next frame (TerrainNode.h:118) */
    static float nextGroundHeightAtCamera;

    /* root quad [-size, size]^2 like <terrainNode size= zmin= zmax= splitFactor= maxLevel=> */
    TerrainNode(float size, float zmin, float zmax, float splitFactor, int maxLevel);
    virtual ~TerrainNode();

    float getSplitFactor() const { return splitFactor; }
    float getSplitDistance() const { return splitDist; }
    float getDistFactor() const { return distFactor; }
    void getLocalCamera(double *x, double *y, double *z) const { *x = camx; *y = camy; *z = camz; }
    /* TerrainNode.cpp:92-97 */
    float getCameraDist(double xmin, double xmax, double ymin, double ymax, double zmin, double zmax) const;
    /* splitFactor * viewportWidth / 1024 * tan(40 deg) / tan(fov / 2), at least 1.1 (TerrainNode.cpp:129-132) */
    static float splitDistance(float splitFactor, float viewportWidth, float fovRadians);
    /* one frame: camera in local space (flat terrains: distFactor 1), then the recursive quad update */
    void update(double camx, double camy, double camz, float splitDist, float distFactor = 1.0f);

private:
    float splitFactor;
    float splitDist;
    float distFactor;
    double camx, camy, camz;
};

}  // namespace proland

#endif
