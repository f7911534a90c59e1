#include "proland/terrain/TileSampler.h"

#include <cassert>

namespace proland
{

TileSampler::Tree::Tree(Tree *parent) : newTree(true), needTile(false), parent(parent), t(NULL)
{
    children[0] = children[1] = children[2] = children[3] = NULL;
}

TileSampler::Tree::~Tree()
{
}

TileSampler::TileSampler(const std::string &name, ptr<TileProducer> producer) :
    Object("TileSampler"), name(name), producer(producer), root(NULL), storeLeaf(true), storeParent(true), async(false),
    held(0)
{
}

TileSampler::~TileSampler()
{
    release();
}

void TileSampler::release()
{
    if (root != NULL) {
        recursiveDelete(root);
        root = NULL;
    }
}

void TileSampler::setAsynchronous(bool v)
{
    async = v;
    assert(!async || storeParent);
}

void TileSampler::recursiveDelete(Tree *t)
{
    if (t->t != NULL) {
        producer->putTile(t->t);
        t->t = NULL;
        --held;
    }
    if (t->children[0] != NULL) {
        for (int i = 0; i < 4; ++i) {
            recursiveDelete(t->children[i]);
        }
    }
    delete t;
}

bool TileSampler::needTile(ptr<TerrainQuad> q)
{
    bool need = storeLeaf;
    if (!storeParent && q->children[0] != NULL && producer->hasChildren(q->level, q->tx, q->ty)) {
        need = false;
    }
    return need;
}

ptr<TaskGraph> TileSampler::update(ptr<TerrainQuad> q, unsigned int frameNumber)
{
    (void) frameNumber;
    ptr<TaskGraph> result = new TaskGraph();
    if (!async && storeLeaf && root != NULL) {
        /* spare capacity goes to the children of the new leaves (TileSampler.cpp:312-315) */
        int prefetchCount = producer->getCache()->getUnusedTiles() + producer->getCache()->getStorage()->getFreeSlots();
        prefetch(root, q, prefetchCount);
    }
    putTiles(&root, q);
    getTiles(NULL, &root, q, result);
    return result;
}

void TileSampler::putTiles(Tree **t, ptr<TerrainQuad> q)
{
    if (*t == NULL) {
        return;
    }
    assert(producer->hasTile(q->level, q->tx, q->ty));
    (*t)->needTile = needTile(q);
    if (!(*t)->needTile && (*t)->t != NULL) {
        producer->putTile((*t)->t);
        (*t)->t = NULL;
        --held;
    }
    if (q->children[0] == NULL) {
        if ((*t)->children[0] != NULL) {
            for (int i = 0; i < 4; ++i) {
                recursiveDelete((*t)->children[i]);
                (*t)->children[i] = NULL;
            }
        }
    } else if (producer->hasChildren(q->level, q->tx, q->ty)) {
        for (int i = 0; i < 4; ++i) {
            putTiles(&((*t)->children[i]), q->children[i]);
        }
    }
}

void TileSampler::getTiles(Tree *parent, Tree **t, ptr<TerrainQuad> q, ptr<TaskGraph> result)
{
    if (*t == NULL) {
        *t = new Tree(parent);
        (*t)->needTile = needTile(q);
        if (q->level == 0 && producer->getRootQuadSize() == 0.0f) {
            producer->setRootQuadSize((float) q->l);
        }
    }
    assert(producer->hasTile(q->level, q->tx, q->ty));

    if ((*t)->needTile) {
        if ((*t)->t == NULL) {
            if (async && q->level > 0) {
                (*t)->t = producer->findTile(q->level, q->tx, q->ty, true);
                if ((*t)->t == NULL) {
                    if (q->isLeaf()) {
                        producer->prefetchTile(q->level, q->tx, q->ty);
                    }
                } else {
                    (*t)->t = producer->getTile(q->level, q->tx, q->ty, 0);
                    assert((*t)->t != NULL);
                    ++held;
                }
            } else {
                (*t)->t = producer->getTile(q->level, q->tx, q->ty, 0);
                if ((*t)->t == NULL) {
                    if (Logger::ERROR_LOGGER != NULL) {
                        Logger::ERROR_LOGGER->log("TERRAIN", "Insufficient tile cache size for '" + name + "' uniform");
                    }
                    throw CacheFullError("Insufficient tile cache size for '" + name + "' uniform");
                }
                ++held;
            }
        }
        if ((*t)->t != NULL && !(*t)->t->task->isDone()) {
            result->addTask((*t)->t->task);
        }
    }

    if (q->children[0] != NULL && producer->hasChildren(q->level, q->tx, q->ty)) {
        for (int i = 0; i < 4; ++i) {
            getTiles(*t, &((*t)->children[i]), q->children[i], result);
        }
    }
}

void TileSampler::prefetch(Tree *t, ptr<TerrainQuad> q, int &prefetchCount)
{
    if (t->children[0] == NULL) {
        if (t->newTree && q != NULL && producer->hasChildren(q->level, q->tx, q->ty)) {
            for (int c = 0; c < 4 && prefetchCount > 0; ++c) {
                if (producer->prefetchTile(q->level + 1, 2 * q->tx + (c & 1), 2 * q->ty + (c >> 1))) {
                    --prefetchCount;
                }
            }
        }
    } else {
        for (int i = 0; i < 4; ++i) {
            prefetch(t->children[i], q == NULL ? ptr<TerrainQuad>() : q->children[i], prefetchCount);
        }
    }
    t->newTree = false;
}

}  // namespace proland
