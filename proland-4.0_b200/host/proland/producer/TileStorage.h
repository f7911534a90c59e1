/*
 * TileStorage -- fixed-capacity pool of tile slots.
 * Mirrors producer/TileStorage.h:45-222 / TileStorage.cpp:37-122 of the reference
 * (same public members); the free list is a deque used FIFO like the reference's
 * std::list (newSlot takes the front, deleteSlot appends).
 */
#ifndef PROLAND_B200_TILE_STORAGE_H
#define PROLAND_B200_TILE_STORAGE_H

#include <deque>
#include <mutex>
#include <utility>

#include "ork/ork_lite.h"

using namespace ork;

namespace proland
{

PROLAND_API class TileStorage : public Object
{
public:
    class Slot
    {
    public:
        /* (producer id, (level, (tx, ty))) of the tile stored here */
        std::pair<int, std::pair<int, std::pair<int, int> > > id;
        /* the task that is producing / has produced the content of this slot */
        void *producerTask;

        Slot(TileStorage *owner);
        virtual ~Slot();
        TileStorage *getOwner();
        void lock(bool lock);

    private:
        TileStorage *owner;
        std::mutex mutex;
    };

    TileStorage(int tileSize, int capacity);
    virtual ~TileStorage();

    Slot *newSlot();
    void deleteSlot(Slot *t);
    int getTileSize();
    int getCapacity();
    int getFreeSlots();

protected:
    int tileSize;
    int capacity;
    std::deque<Slot *> freeSlots;

    TileStorage();
    void init(int tileSize, int capacity);
};

}  // namespace proland

#endif
