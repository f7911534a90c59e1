#include "proland/terrain/TerrainQuad.h"

#include <algorithm>
#include <cassert>
#include <cmath>

namespace proland
{

float TerrainNode::groundHeightAtCamera = 0.0f;
float TerrainNode::nextGroundHeightAtCamera = 0.0f;

TerrainQuad::TerrainQuad(TerrainNode *owner, const TerrainQuad *parent, int tx, int ty, double ox, double oy, double l,
                         float zmin, float zmax) :
    Object("TerrainQuad"), parent(parent), level(parent == NULL ? 0 : parent->level + 1), tx(tx), ty(ty), ox(ox), oy(oy),
    l(l), zmin(zmin), zmax(zmax), owner(owner)
{
}

TerrainQuad::~TerrainQuad()
{
}

TerrainNode *TerrainQuad::getOwner()
{
    return owner;
}

bool TerrainQuad::isLeaf() const
{
    return children[0] == NULL;
}

int TerrainQuad::getSize() const
{
    if (isLeaf()) {
        return 1;
    }
    return 1 + children[0]->getSize() + children[1]->getSize() + children[2]->getSize() + children[3]->getSize();
}

int TerrainQuad::getDepth() const
{
    if (isLeaf()) {
        return level;
    }
    return std::max(std::max(children[0]->getDepth(), children[1]->getDepth()),
                    std::max(children[2]->getDepth(), children[3]->getDepth()));
}

void TerrainQuad::update()
{
    const double ground = TerrainNode::groundHeightAtCamera;
    const float dist = owner->getCameraDist(ox, ox + l, oy, oy + l, std::min(0.0, ground), std::max(0.0, ground));

    if (dist < l * owner->getSplitDistance() && level < owner->maxLevel) {
        if (isLeaf()) {
            subdivide();
        }
        /* nearest child first (TerrainQuad.cpp:110-150) */
        double cx, cy, cz;
        owner->getLocalCamera(&cx, &cy, &cz);
        const double mx = ox + l / 2.0, my = oy + l / 2.0;
        static const int orders[4][4] = { { 0, 1, 2, 3 }, { 1, 0, 3, 2 }, { 2, 0, 3, 1 }, { 3, 1, 2, 0 } };
        const int *order = orders[(cy < my ? 0 : 2) + (cx < mx ? 0 : 1)];
        for (int i = 0; i < 4; ++i) {
            children[order[i]]->update();
        }
    } else if (!isLeaf()) {
        for (int i = 0; i < 4; ++i) {
            children[i] = NULL;
        }
    }
}

void TerrainQuad::subdivide()
{
    const float hl = (float) l / 2.0f;
    children[0] = new TerrainQuad(owner, this, 2 * tx, 2 * ty, ox, oy, hl, zmin, zmax);
    children[1] = new TerrainQuad(owner, this, 2 * tx + 1, 2 * ty, ox + hl, oy, hl, zmin, zmax);
    children[2] = new TerrainQuad(owner, this, 2 * tx, 2 * ty + 1, ox, oy + hl, hl, zmin, zmax);
    children[3] = new TerrainQuad(owner, this, 2 * tx + 1, 2 * ty + 1, ox + hl, oy + hl, hl, zmin, zmax);
}

TerrainNode::TerrainNode(float size, float zmin, float zmax, float splitFactor, int maxLevel) :
    Object("TerrainNode"), maxLevel(maxLevel), splitFactor(splitFactor), splitDist(1.1f), distFactor(1.0f), camx(0.0),
    camy(0.0), camz(0.0)
{
    root = new TerrainQuad(this, NULL, 0, 0, -size, -size, 2.0 * size, zmin, zmax);
}

TerrainNode::~TerrainNode()
{
}

float TerrainNode::getCameraDist(double xmin, double xmax, double ymin, double ymax, double zmin, double zmax) const
{
    (void) zmin;
    return (float) std::max(std::abs(camz - zmax) / distFactor,
                            std::max(std::min(std::abs(camx - xmin), std::abs(camx - xmax)),
                                     std::min(std::abs(camy - ymin), std::abs(camy - ymax))));
}

float TerrainNode::splitDistance(float splitFactor, float viewportWidth, float fovRadians)
{
    float d = splitFactor * viewportWidth / 1024.0f * tanf(40.0f / 180.0f * (float) M_PI) / tanf(fovRadians / 2.0f);
    if (d < 1.1f || !std::isfinite(d)) {
        d = 1.1f;
    }
    return d;
}

void TerrainNode::update(double x, double y, double z, float splitDistance, float distanceFactor)
{
    assert(splitDistance > 1.0f);
    camx = x;
    camy = y;
    camz = z;
    splitDist = splitDistance;
    distFactor = distanceFactor;
    root->update();
}

}  // namespace proland
