#include "proland/producer/TileStorage.h"

namespace proland
{

TileStorage::Slot::Slot(TileStorage *owner) : producerTask(NULL), owner(owner)
{
    id = std::make_pair(-1, std::make_pair(-1, std::make_pair(-1, -1)));
}

TileStorage::Slot::~Slot()
{
}

TileStorage *TileStorage::Slot::getOwner()
{
    return owner;
}

void TileStorage::Slot::lock(bool lock)
{
    if (lock) {
        mutex.lock();
    } else {
        mutex.unlock();
    }
}

TileStorage::TileStorage(int tileSize, int capacity) : Object("TileStorage")
{
    init(tileSize, capacity);
}

TileStorage::TileStorage() : Object("TileStorage"), tileSize(0), capacity(0)
{
}

void TileStorage::init(int tileSize, int capacity)
{
    this->tileSize = tileSize;
    this->capacity = capacity;
}

TileStorage::~TileStorage()
{
    for (size_t i = 0; i < freeSlots.size(); ++i) {
        delete freeSlots[i];
    }
    freeSlots.clear();
}

TileStorage::Slot *TileStorage::newSlot()
{
    if (freeSlots.empty()) {
        return NULL;
    }
    Slot *s = freeSlots.front();
    freeSlots.pop_front();
    return s;
}

void TileStorage::deleteSlot(TileStorage::Slot *t)
{
    freeSlots.push_back(t);
}

int TileStorage::getTileSize()
{
    return tileSize;
}

int TileStorage::getCapacity()
{
    return capacity;
}

int TileStorage::getFreeSlots()
{
    return (int) freeSlots.size();
}

}  // namespace proland
