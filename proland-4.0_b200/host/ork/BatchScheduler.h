/*
 * BatchScheduler -- runs task graphs in waves of mutually independent tasks.
 *
 * Takes the place of Ork's MultithreadScheduler (`<multithreadScheduler
 * nthreads= fps= prefetchRate= prefetchQueue=>`, fractalterrain.xml:27,
 * earth-srtm-async.xml:27) for the tile-production path.  Ork runs each task as
 * begin / run / end on the GL thread, one draw call per tile; here every wave
 * (all tasks whose dependencies are done) is executed between
 * DeviceContext::beginBatchAll / endBatchAll, so that the producers only QUEUE
 * their tiles during run() and each producer launches ONE kernel per wave.  For a
 * quadtree request that is one elevation and one normal launch per level.
 *
 * Prefetch (TileProducer::prefetchTile -> schedule): at most `prefetchRate`
 * queued tasks are added to each run(), out of a queue of `prefetchQueue` entries.
 */
#ifndef PROLAND_B200_BATCH_SCHEDULER_H
#define PROLAND_B200_BATCH_SCHEDULER_H

#include <deque>

#include "ork/ork_lite.h"

namespace ork
{

class BatchScheduler : public Scheduler
{
public:
    /* called around every wave; the device layer installs its batch brackets here */
    typedef void (*WaveHook)(bool begin);
    static void setWaveHook(WaveHook hook);

    BatchScheduler(int prefetchRate = 0, int prefetchQueue = 0);
    virtual ~BatchScheduler();

    virtual bool supportsPrefetch(bool gpuTasks);
    virtual void schedule(ptr<Task> task);
    virtual void reschedule(ptr<Task> task, Task::reason r, unsigned int deadline);
    virtual void run(ptr<Task> task);
    /* forget the queued prefetch tasks (before the caches they belong to go away) */
    void clear() { prefetch.clear(); }

    /* the date given to the tasks of the current / last run (Ork: the frame number) */
    unsigned int getFrame() const { return frame; }
    unsigned long getWaveCount() const { return waves; }
    unsigned long getTaskCount() const { return executed; }
    int getQueuedPrefetchCount() const { return (int) prefetch.size(); }

private:
    int prefetchRate;
    int prefetchQueue;
    std::deque<ptr<Task> > prefetch;
    unsigned int frame;
    unsigned long waves, executed;
    size_t lastNodes;      /* size of the last flattened view: the next one reserves as much */
};

}  // namespace ork

#endif
