#include "ork/BatchScheduler.h"

#include <algorithm>
#include <atomic>
#include <map>
#include <unordered_map>
#include <unordered_set>
#include <stdexcept>

namespace ork
{

static BatchScheduler::WaveHook g_hook = NULL;

void BatchScheduler::setWaveHook(WaveHook hook)
{
    g_hook = hook;
}

BatchScheduler::BatchScheduler(int prefetchRate, int prefetchQueue) :
    Scheduler("BatchScheduler"), prefetchRate(prefetchRate), prefetchQueue(prefetchQueue), frame(0), waves(0), executed(0), lastNodes(0)
{
}

BatchScheduler::~BatchScheduler()
{
}

bool BatchScheduler::supportsPrefetch(bool gpuTasks)
{
    (void) gpuTasks;
    return prefetchQueue > 0 && prefetchRate > 0;
}

void BatchScheduler::schedule(ptr<Task> task)
{
    if ((int) prefetch.size() >= prefetchQueue && !prefetch.empty()) {
        prefetch.pop_front();   /* the oldest request is the least likely to still matter */
    }
    prefetch.push_back(task);
}

void BatchScheduler::reschedule(ptr<Task> task, Task::reason r, unsigned int deadline)
{
    task->setIsDone(false, 0, r);
    task->setDeadline(deadline);
}

namespace
{

struct Node
{
    Task *task;
    void *context;              /* Task::getContext, read once per flattened view (the sort key of a wave) */
    std::vector<Task *> deps;   /* tasks or graphs this task waits for, over all graphs that hold it */
};

/* stamps handed to Task::schedViewStamp / schedMemoStamp: never 0, never reused */
static std::atomic<unsigned long long> g_stamp(0);      /* schedulers of different threads draw from one sequence */

/* the primitive tasks below the roots, in the order the graphs list them.  A task knows its slot in the current view
 * (Task::schedSlot, valid while Task::schedViewStamp equals the view's stamp): no look-up table */
struct Flat
{
    std::vector<Node> nodes;
    unsigned long long stamp;       /* of this view */
    unsigned long long madeAt;      /* TaskGraph::editCount() when the view was made */

    Flat() : stamp(++g_stamp), madeAt(~0ull) {}

    Node &operator[](Task *t)
    {
        if (t->schedViewStamp == stamp) {
            return nodes[t->schedSlot];
        }
        t->schedViewStamp = stamp;
        t->schedSlot = nodes.size();
        nodes.push_back(Node());
        nodes.back().task = t;
        nodes.back().context = NULL;
        return nodes.back();
    }

    /* first visit of a graph in this view? */
    bool enter(Task *g)
    {
        if (g->schedViewStamp == stamp) {
            return false;
        }
        g->schedViewStamp = stamp;
        return true;
    }
};

/* primitive tasks below `t`, with their dependencies */
void flatten(Task *t, Flat &nodes)
{
    if (!t->isTaskGraph()) {
        Node &n = nodes[t];
        n.task = t;
        return;
    }
    if (!nodes.enter(t)) {
        return;
    }
    TaskGraph *g = static_cast<TaskGraph *>(t);
    const TaskGraph::TaskSet &ts = g->taskSet();
    for (TaskGraph::TaskSet::const_iterator i = ts.begin(); i != ts.end(); ++i) {
        Task *c = i->get();
        flatten(c, nodes);
        const TaskGraph::TaskSet *deps = g->dependenciesOf(c);
        if (deps == NULL) {
            continue;
        }
        if (!c->isTaskGraph()) {
            Node &n = nodes[c];
            for (TaskGraph::TaskSet::const_iterator d = deps->begin(); d != deps->end(); ++d) n.deps.push_back(d->get());
        } else {
            /* a sub-graph waiting for something: its first tasks wait */
            TaskGraph::TaskIterator f = static_cast<TaskGraph *>(c)->getFirstTasks();
            while (f.hasNext()) {
                Task *ft = f.next().get();
                if (ft->isTaskGraph()) continue;
                Node &n = nodes[ft];
                n.task = ft;
                for (TaskGraph::TaskSet::const_iterator d = deps->begin(); d != deps->end(); ++d) n.deps.push_back(d->get());
            }
        }
    }
}

bool byContext(const Node *a, const Node *b)
{
    return a->context != b->context ? a->context < b->context : a->task < b->task;
}

/* is a dependency (a task or a graph) done?  A graph is done when all its tasks are (TaskGraph::isDone walks them,
 * sub-graphs included: a tile's graph holds its parent's, which holds its parent's ...); within one wave nothing
 * completes, so the answer per graph is remembered for the wave */
struct DoneMemo
{
    unsigned long long stamp;

    DoneMemo() : stamp(++g_stamp) {}

    bool operator()(Task *t)
    {
        if (!t->isTaskGraph()) {
            return t->isDone();
        }
        if (t->schedMemoStamp == stamp) {
            return t->schedMemo;
        }
        bool done = true;
        const TaskGraph::TaskSet &ts = static_cast<TaskGraph *>(t)->taskSet();
        for (TaskGraph::TaskSet::const_iterator c = ts.begin(); c != ts.end() && done; ++c) {
            done = (*this)(c->get());
        }
        t->schedMemoStamp = stamp;
        t->schedMemo = done;
        return done;
    }
};

}  // namespace

static void rebuild(const std::vector<ptr<Task> > &roots, Flat &nodes)
{
    const size_t hint = nodes.nodes.size();
    nodes = Flat();
    nodes.nodes.reserve(hint);
    for (size_t i = 0; i < roots.size(); ++i) {
        flatten(roots[i].get(), nodes);
    }
    nodes.madeAt = TaskGraph::editCount();
}

/* Task::init of every primitive task below the roots that is not in `initialized` yet; initialising a task may add
 * tasks (CreateTile::start acquires the tiles it is made from): the view is rebuilt and the pass repeated while the
 * graphs keep changing (TaskGraph::editCount).  Leaves `nodes` current. */
static void initAll(const std::vector<ptr<Task> > &roots, std::set<Task *> &initialized, Flat &nodes)
{
    for (;;) {
        if (nodes.madeAt != TaskGraph::editCount()) {
            rebuild(roots, nodes);
        }
        /* keep the tasks alive: an init may drop the last other reference to a sibling */
        std::vector<ptr<Task> > todo;
        for (std::vector<Node>::iterator n = nodes.nodes.begin(); n != nodes.nodes.end(); ++n) {
            if (initialized.insert(n->task).second) {
                todo.push_back(n->task);
            }
        }
        for (size_t i = 0; i < todo.size(); ++i) {
            todo[i]->init(initialized);
        }
        if (todo.empty() || nodes.madeAt == TaskGraph::editCount()) {
            break;
        }
    }
}

void BatchScheduler::run(ptr<Task> task)
{
    ++frame;
    std::vector<ptr<Task> > roots;
    if (task != NULL) {
        roots.push_back(task);
    }
    for (int k = 0; k < prefetchRate && !prefetch.empty(); ++k) {
        roots.push_back(prefetch.front());
        prefetch.pop_front();
    }

    /* Task::init of every primitive task below the roots, once.  TaskGraph::init would walk the nested
     * per-tile graphs recursively -- a tile's graph holds its parent's, which holds its parent's ...: every
     * frame re-walked each ancestor chain once per descendant (60 % of the host time of a fly-through).  The
     * flattened view visits every graph once; initialising a task may add tasks (CreateTile::start acquires
     * the tiles it is made from), so flatten again until nothing new appears. */
    std::set<Task *> initialized;
    Flat nodes;
    nodes.nodes.reserve(lastNodes);
    initAll(roots, initialized, nodes);

    for (;;) {
        /* the flattened view is rebuilt only when some graph was edited since it was made (restarting a task
         * re-acquires its inputs and edits its graph; finishing one only releases tiles): TaskGraph::editCount */
        if (nodes.madeAt != TaskGraph::editCount()) {
            rebuild(roots, nodes);
        }
        /* done tasks that completed before one of their inputs changed are stale */
        bool restarted = false;
        for (std::vector<Node>::iterator n = nodes.nodes.begin(); n != nodes.nodes.end(); ++n) {
            Task *t = n->task;
            if (!t->isDone()) continue;
            /* a task completed in this run carries this run's date: none of its inputs can have changed LATER */
            if (t->getCompletionDate() == frame) continue;
            for (size_t d = 0; d < n->deps.size(); ++d) {
                if (n->deps[d]->getChangeDate() > t->getCompletionDate()) {
                    t->setIsDone(false, 0, Task::DATA_CHANGED);
                    initialized.erase(t);
                    restarted = true;
                    break;
                }
            }
        }
        if (restarted) {
            /* restarted tasks re-acquire their input tiles (CreateTile::init -> start) */
            initAll(roots, initialized, nodes);
            continue;
        }

        std::vector<Node *> readyNodes;
        size_t open = 0;
        DoneMemo isDone;
        for (std::vector<Node>::iterator n = nodes.nodes.begin(); n != nodes.nodes.end(); ++n) {
            if (n->task->isDone()) continue;
            ++open;
            bool ok = true;
            for (size_t d = 0; d < n->deps.size() && ok; ++d) {
                ok = isDone(n->deps[d]);
            }
            if (ok) {
                if (n->context == NULL) n->context = n->task->getContext();
                readyNodes.push_back(&*n);
            }
        }
        if (open == 0) {
            break;
        }
        if (readyNodes.empty()) {
            if (Logger::ERROR_LOGGER != NULL) {
                Logger::ERROR_LOGGER->logf("SCHEDULER", "%d tasks left but none can run (dependency cycle)", (int) open);
            }
            throw std::logic_error("BatchScheduler: dependency cycle");
        }
        std::sort(readyNodes.begin(), readyNodes.end(), byContext);
        std::vector<Task *> ready(readyNodes.size());
        for (size_t i = 0; i < readyNodes.size(); ++i) ready[i] = readyNodes[i]->task;

        /* keep the tasks alive while they run */
        std::vector<ptr<Task> > hold(ready.begin(), ready.end());
        std::vector<bool> changes(ready.size(), true);
        if (g_hook) g_hook(true);
        try {
            for (size_t i = 0; i < ready.size(); ++i) {
                ready[i]->begin();
                changes[i] = ready[i]->run();
                ready[i]->end();
            }
        } catch (...) {
            if (g_hook) g_hook(false);
            throw;
        }
        if (g_hook) g_hook(false);
        for (size_t i = 0; i < ready.size(); ++i) {
            ready[i]->setIsDone(true, frame, changes[i] ? Task::DATA_CHANGED : Task::DATA_NEEDED);
        }
        ++waves;
        executed += ready.size();
    }
    lastNodes = nodes.nodes.size();
    for (size_t i = 0; i < roots.size(); ++i) {
        if (roots[i]->isTaskGraph()) {
            /* a graph has no run(); record its completion */
            roots[i]->Task::setIsDone(true, frame, Task::DATA_NEEDED);
        }
    }
}

}  // namespace ork
