/*
 * ork_lite.h -- the few Ork types the tile-production path of Proland leans on,
 * restated without OpenGL: a reference-counted Object, ptr<T>, Logger, Task,
 * TaskGraph and the Scheduler interface.
 *
 * Ork itself is not part of the reference tree; the contracts below
 * are the ones the reference's producer code relies on:
 *   - ptr<T>/Object: intrusive counting, doRelease() hook
 *     (used by CreateTileTaskGraph, producer/TileProducer.cpp:281-301)
 *   - Task: isDone/setIsDone(done, t, reason), init/begin/run/end, getContext,
 *     deadline (producer/TileProducer.cpp:44-247)
 *   - TaskGraph: addTask/removeTask/addDependency/clearDependencies/
 *     getLastTasks/removeAndGetDependencies (TileProducer.cpp:155-176,281-316)
 *   - Scheduler: run/schedule/reschedule/supportsPrefetch
 *     (producer/TileCache.cpp:228-232, TileProducer.cpp:521-532)
 */
#ifndef PROLAND_B200_ORK_LITE_H
#define PROLAND_B200_ORK_LITE_H

#include <atomic>
#include <cstddef>
#include <map>
#include <set>
#include <string>
#include <typeinfo>
#include <unordered_set>
#include <vector>

#ifndef PROLAND_API
#define PROLAND_API
#endif

namespace ork
{

/* ork/math/vec4.h: what the preprocessing interface needs of it (InputMap::getValue returns one) */
template <typename T> struct vec4
{
    T x, y, z, w;
    vec4() : x(0), y(0), z(0), w(0) {}
    vec4(T x, T y, T z, T w) : x(x), y(y), z(z), w(w) {}
};
typedef vec4<float> vec4f;

class Object
{
public:
    explicit Object(const char *type) : type(type), references(0) {}
    virtual ~Object() {}
    const char *getClass() const { return type; }
    void acquire() { references.fetch_add(1, std::memory_order_relaxed); }
    void release()
    {
        if (references.fetch_sub(1, std::memory_order_acq_rel) == 1) {
            doRelease();
        }
    }
    int getReferences() const { return references.load(); }

protected:
    /* called when the last ptr<> goes away; the default deletes the object */
    virtual void doRelease() { delete this; }

private:
    const char *type;
    std::atomic<int> references;
};

template <class T>
class ptr
{
public:
    ptr() : target(NULL) {}
    ptr(T *t) : target(t) { if (target) target->acquire(); }
    ptr(const ptr<T> &p) : target(p.target) { if (target) target->acquire(); }
    template <class U>
    ptr(const ptr<U> &p) : target(p.get()) { if (target) target->acquire(); }
    ~ptr() { if (target) target->release(); }

    ptr<T> &operator=(const ptr<T> &p) { reset(p.target); return *this; }
    ptr<T> &operator=(T *t) { reset(t); return *this; }
    template <class U>
    ptr<T> &operator=(const ptr<U> &p) { reset(p.get()); return *this; }

    T *operator->() const { return target; }
    T &operator*() const { return *target; }
    T *get() const { return target; }
    template <class U>
    ptr<U> cast() const { return ptr<U>(dynamic_cast<U *>(target)); }

    bool operator==(const ptr<T> &p) const { return target == p.target; }
    bool operator!=(const ptr<T> &p) const { return target != p.target; }
    bool operator==(const T *t) const { return target == t; }
    bool operator!=(const T *t) const { return target != t; }
    bool operator<(const ptr<T> &p) const { return target < p.target; }

private:
    void reset(T *t)
    {
        if (t) t->acquire();
        T *old = target;
        target = t;
        if (old) old->release();
    }
    T *target;
};

/* Ork's loggers: a topic and a message.  DEBUG is NULL unless a test or a tool installs one (the
 * reference's producers log "Elevation tile <id> <level> <tx> <ty>" there, ElevationProducer.cpp:282-286). */
class Logger
{
public:
    explicit Logger(const char *level) : level(level), lines(0), echo(true) {}
    virtual ~Logger() {}
    virtual void log(const std::string &topic, const std::string &msg);
    void logf(const char *topic, const char *fmt, ...);
    unsigned long getLineCount() const { return lines; }
    void setEcho(bool e) { echo = e; }
    const std::string &getLastLine() const { return last; }

    static Logger *DEBUG_LOGGER;
    static Logger *INFO_LOGGER;
    static Logger *WARNING_LOGGER;
    static Logger *ERROR_LOGGER;

private:
    const char *level;
    unsigned long lines;
    bool echo;
    std::string last;
};

class TaskGraph;

class Task : public Object
{
public:
    enum reason { STRUCTURE_CHANGED, DATA_CHANGED, DATA_NEEDED };

    Task(const char *type, bool gpuTask, unsigned int deadline);
    virtual ~Task();

    /* tasks with equal contexts can run back to back without a context switch; here: can share a batch */
    virtual void *getContext() const { return NULL; }
    bool isGpuTask() const { return gpuTask; }
    unsigned int getDeadline() const { return deadline; }
    virtual void setDeadline(unsigned int d) { if (d < deadline) deadline = d; }

    virtual bool isDone() { return done; }
    /* done -> completion date t; !done -> the task must run again, for the given reason */
    virtual void setIsDone(bool done, unsigned int t, reason r = DATA_NEEDED);
    virtual unsigned int getCompletionDate() { return completionDate; }
    /* the last date this task ran because its data had CHANGED (invalidation), not merely because the
     * data was needed again: tasks that depend on it and completed earlier are stale */
    virtual unsigned int getChangeDate() { return changeDate; }
    reason getLastReason() const { return lastReason; }

    /* called once per Scheduler::run before any task of the graph runs */
    virtual void init(std::set<Task *> &initialized) { (void) initialized; }
    virtual void begin() {}
    virtual bool run() { return true; }
    virtual void end() {}
    virtual const std::type_info *getTypeInfo() { return &typeid(*this); }

    virtual bool isTaskGraph() const { return false; }

    /* scratch of the scheduler that is flattening the graphs this task sits in (one scheduler thread at a time):
     * the task's slot in the current flattened view and a per-wave memo, each valid for one stamp -- instead of hash
     * tables keyed by the task's address, rebuilt every view */
    mutable unsigned long long schedViewStamp;
    mutable size_t schedSlot;
    mutable unsigned long long schedMemoStamp;
    mutable bool schedMemo;

protected:
    bool done;
    unsigned int completionDate;
    unsigned int changeDate;
    reason lastReason;

private:
    bool gpuTask;
    unsigned int deadline;
};

/* A set of tasks with dependencies "src needs dst".  A graph is done when all its tasks are. */
class TaskGraph : public Task
{
public:
    /* unique tasks in insertion order.  Ork keeps std::set / std::map here; a tile's graph holds two or three tasks and one
     * or two dependencies, and every container node was a heap allocation per tile: flat vectors, searched linearly, with a
     * hash index only for graphs that grow large (a frame's root graph) */
    typedef std::vector<ptr<Task> > TaskSet;

    /* Ork's iterator flavour: hasNext()/next() over a snapshot */
    class TaskIterator
    {
    public:
        TaskIterator() : pos(0) {}
        explicit TaskIterator(const std::vector<ptr<Task> > &v) : tasks(v), pos(0) {}
        bool hasNext() const { return pos < tasks.size(); }
        ptr<Task> next() { return tasks[pos++]; }
    private:
        std::vector<ptr<Task> > tasks;
        size_t pos;
    };

    TaskGraph();
    virtual ~TaskGraph();

    virtual bool isTaskGraph() const { return true; }
    virtual bool isDone();
    virtual void setIsDone(bool done, unsigned int t, reason r = DATA_NEEDED);
    virtual unsigned int getCompletionDate();
    virtual unsigned int getChangeDate();
    virtual void init(std::set<Task *> &initialized);

    bool isEmpty() const { return tasks.empty(); }
    void addTask(ptr<Task> t);
    void removeTask(ptr<Task> t);
    /* src can only run once dst is done */
    void addDependency(ptr<Task> src, ptr<Task> dst);
    void removeDependency(ptr<Task> src, ptr<Task> dst);
    void removeAndGetDependencies(ptr<Task> src, TaskSet &deletedDependencies);
    void clearDependencies();

    TaskIterator getAllTasks() const;
    TaskIterator getFirstTasks() const;          /* tasks that need no other task of this graph */
    TaskIterator getLastTasks() const;           /* tasks no other task of this graph needs */
    TaskIterator getDependencies(ptr<Task> t) const;
    TaskIterator getInverseDependencies(ptr<Task> t) const;

    /* direct access for schedulers */
    const TaskSet &taskSet() const { return tasks; }
    /* counts the structural edits (tasks or dependencies added / removed) of ALL graphs: a scheduler that keeps a
     * flattened view across waves rebuilds it only when this moved */
    static unsigned long long editCount() { return edits.load(std::memory_order_relaxed); }
    const TaskSet *dependenciesOf(Task *t) const;

protected:
    /* drop every strong reference this graph holds (Ork: TaskGraph::cleanup) */
    void cleanup();

private:
    struct Needs { Task *src; TaskSet dst; };                    /* src -> what it needs */
    struct NeededBy { Task *dst; std::vector<Task *> src; };      /* dst -> who needs it */
    static std::atomic<unsigned long long> edits;
    TaskSet tasks;
    std::unordered_set<Task *> *index;                            /* of `tasks`, once there are more than kIndexAbove */
    std::vector<Needs> dependencies;
    std::vector<NeededBy> inverse;
    enum { kIndexAbove = 16 };

    bool contains(Task *t) const;
    Needs *needsOf(Task *src);
    const Needs *needsOf(Task *src) const;
    NeededBy *neededBy(Task *dst);
    const NeededBy *neededBy(Task *dst) const;
};

class Scheduler : public Object
{
public:
    explicit Scheduler(const char *type) : Object(type) {}
    virtual ~Scheduler() {}
    virtual bool supportsPrefetch(bool gpuTasks) = 0;
    /* queue a task for a later run() (prefetch) */
    virtual void schedule(ptr<Task> task) = 0;
    /* the task must be (re)executed: setIsDone(false, 0, r) + bookkeeping */
    virtual void reschedule(ptr<Task> task, Task::reason r, unsigned int deadline) = 0;
    /* execute the task (graph) and, budget permitting, queued prefetch tasks */
    virtual void run(ptr<Task> task) = 0;
};

}  // namespace ork

#endif
