#include "proland/terrain/TileSamplerZ.h"

#include <cassert>
#include <cmath>

#include "proland/producer/DeviceContext.h"

namespace proland
{

std::map<GPUTileStorage *, TileSamplerZ::State *> TileSamplerZ::states;

TileSamplerZ::TreeZ::TreeZ(Tree *parent, ptr<TerrainQuad> q) : Tree(parent), q(q), readback(false), readbackDate(0)
{
}

bool TileSamplerZ::TreeZSort::operator()(const TreeZ *x, const TreeZ *y) const
{
    const int xl = x->q->level, yl = y->q->level;
    return xl == yl ? x < y : xl < yl;
}

TileSamplerZ::TileSamplerZ(const std::string &name, ptr<TileProducer> producer) :
    TileSampler(name, producer), state(NULL), cameraQuad(NULL), cameraQuadX(0.0f), cameraQuadY(0.0f), oldCamX(0.0), oldCamY(0.0),
    oldCamZ(0.0)
{
    GPUTileStorage *storage = dynamic_cast<GPUTileStorage *>(producer->getCache()->getStorage().get());
    if (storage == NULL) {
        throw std::invalid_argument("TileSamplerZ needs a producer of GPU tiles");
    }
    std::map<GPUTileStorage *, State *>::iterator i = states.find(storage);
    if (i == states.end()) {
        State *s = new State();
        s->storage = storage;
        s->users = 0;
        s->cameraSlot = NULL;
        s->cameraX = s->cameraY = 0;
        s->lastFrame = 0;
        s->issued = s->applied = 0;
        i = states.insert(std::make_pair(storage, s)).first;
    }
    state = i->second;
    ++state->users;
}

TileSamplerZ::~TileSamplerZ()
{
    release();      /* recursiveDelete takes the trees out of needReadback */
    if (--state->users == 0) {
        collect(0, true);
        states.erase(state->storage);
        delete state;
    }
    state = NULL;
}

void TileSamplerZ::getCounts(unsigned long long out[3]) const
{
    out[0] = state->issued;
    out[1] = state->applied;
    out[2] = state->needReadback.size();
}

/* TileCallback::dataRead (TileSamplerZ.cpp:162-175) for the read-backs old enough (all: every one, discarding nothing) */
void TileSamplerZ::collect(unsigned int frameNumber, bool all)
{
    pl_ctx *ctx = state->storage->getContext()->handle();
    while (!state->pending.empty() && (all || frameNumber - state->pending.front().frame >= (unsigned int) READBACK_DELAY)) {
        State::Pending &p = state->pending.front();
        float values[2 * (MAX_TILES_PER_FRAME + 1)];
        DeviceContext::check(pl_elev_stats_readback_end(ctx, p.ticket, values));
        unsigned int i = 0;
        if (p.camera) {
            TerrainNode::groundHeightAtCamera = TerrainNode::nextGroundHeightAtCamera;
            TerrainNode::nextGroundHeightAtCamera = values[0];
            i = 1;
        }
        for (; i < p.targets.size(); ++i) {
            p.targets[i]->zmin = values[2 * i];
            p.targets[i]->zmax = values[2 * i + 1];
        }
        ++state->applied;
        state->pending.pop_front();
    }
}

ptr<TaskGraph> TileSamplerZ::update(ptr<TerrainQuad> root, unsigned int frameNumber)
{
    ptr<TaskGraph> result = TileSampler::update(root, frameNumber);

    double cx, cy, cz;
    root->getOwner()->getLocalCamera(&cx, &cy, &cz);
    const double moved = std::sqrt((cx - oldCamX) * (cx - oldCamX) + (cy - oldCamY) * (cy - oldCamY) + (cz - oldCamZ) * (cz - oldCamZ));
    if (moved > 0.1 && cameraQuad != NULL && cameraQuad->t != NULL) {
        GPUTileStorage::GPUSlot *gpuTile = dynamic_cast<GPUTileStorage::GPUSlot *>(cameraQuad->t->getData(false));
        if (gpuTile != NULL && state->cameraSlot == NULL) {
            const int border = get()->getBorder();
            const int tileSize = get()->getCache()->getStorage()->getTileSize() - 2 * border;
            const int dx = std::min((int) std::floor(cameraQuadX * tileSize), tileSize - 1);
            const int dy = std::min((int) std::floor(cameraQuadY * tileSize), tileSize - 1);
            assert(border == 2);
            state->cameraSlot = gpuTile;
            state->cameraX = dx + border;
            state->cameraY = dy + border;
            oldCamX = cx;
            oldCamY = cy;
            oldCamZ = cz;
        }
    }
    cameraQuad = NULL;

    if (frameNumber == state->lastFrame) {
        return result;      /* another sampler of this storage has done this frame's read-back */
    }
    collect(frameNumber, false);      /* ReadbackManager::newFrame */
    state->lastFrame = frameNumber;

    std::vector<int32_t> slots;
    State::Pending p;
    p.camera = false;
    int camSlot = -1;
    if (state->cameraSlot != NULL) {
        camSlot = state->cameraSlot->l;
        p.targets.push_back(NULL);
        p.camera = true;
        state->cameraSlot = NULL;
    }
    std::set<TreeZ *, TreeZSort>::iterator i = state->needReadback.begin();
    while (i != state->needReadback.end() && slots.size() + (p.camera ? 1 : 0) < (size_t) MAX_TILES_PER_FRAME) {
        TreeZ *t = *i;
        TileCache::Tile *tile = t->t;
        state->needReadback.erase(i++);
        if (tile != NULL) {
            GPUTileStorage::GPUSlot *gpuTile = dynamic_cast<GPUTileStorage::GPUSlot *>(tile->getData(false));
            if (gpuTile != NULL) {
                slots.push_back(gpuTile->l);
                p.targets.push_back(t->q);
            } else {
                t->readback = false;
            }
        }
    }
    if (slots.empty() && !p.camera) {
        return result;
    }
    /* the tiles may have been queued this very frame by another path: what is read must have been launched */
    state->storage->getContext()->flush();
    DeviceContext::check(pl_elev_zreadback_begin(state->storage->getContext()->handle(), state->storage->getPool(), (int) slots.size(),
                                                 slots.empty() ? NULL : &slots[0], camSlot, state->cameraX, state->cameraY, &p.ticket));
    p.frame = frameNumber;
    state->pending.push_back(p);
    ++state->issued;
    return result;
}

bool TileSamplerZ::needTile(ptr<TerrainQuad> q)
{
    double cx, cy, cz;
    q->getOwner()->getLocalCamera(&cx, &cy, &cz);
    if (cx >= q->ox && cx < q->ox + q->l && cy >= q->oy && cy < q->oy + q->l) {
        return true;
    }
    return TileSampler::needTile(q);
}

void TileSamplerZ::recursiveDelete(Tree *t)
{
    TreeZ *z = static_cast<TreeZ *>(t);
    state->needReadback.erase(z);
    if (cameraQuad == z) {
        cameraQuad = NULL;
    }
    TileSampler::recursiveDelete(t);
}

void TileSamplerZ::getTiles(Tree *parent, Tree **t, ptr<TerrainQuad> q, ptr<TaskGraph> result)
{
    if (*t == NULL) {
        *t = new TreeZ(parent, q);
        (*t)->needTile = needTile(q);
        if (q->level == 0 && get()->getRootQuadSize() == 0.0f) {
            get()->setRootQuadSize((float) q->l);
        }
    }
    TreeZ *z = static_cast<TreeZ *>(*t);
    if (z->t != NULL && z->t->task->isDone() && (!z->readback || z->readbackDate < z->t->task->getCompletionDate())) {
        state->needReadback.insert(z);
        z->readback = true;
        z->readbackDate = z->t->task->getCompletionDate();
    }

    TileSampler::getTiles(parent, t, q, result);

    if (cameraQuad == NULL && z->t != NULL && z->t->task->isDone()) {
        double cx, cy, cz;
        q->getOwner()->getLocalCamera(&cx, &cy, &cz);
        if (cx >= q->ox && cx < q->ox + q->l && cy >= q->oy && cy < q->oy + q->l) {
            cameraQuadX = (float) ((cx - q->ox) / q->l);
            cameraQuadY = (float) ((cy - q->oy) / q->l);
            cameraQuad = z;
        }
    }
}

}  // namespace proland
