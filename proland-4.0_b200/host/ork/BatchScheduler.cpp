#include "ork/BatchScheduler.h"

#include <algorithm>
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
    Scheduler("BatchScheduler"), prefetchRate(prefetchRate), prefetchQueue(prefetchQueue), frame(0), waves(0), executed(0)
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
    std::vector<Task *> deps;   /* tasks or graphs this task waits for, over all graphs that hold it */
};

/* the primitive tasks below the roots, in the order the graphs list them (hashed lookup: the view is
 * rebuilt every wave of every frame) */
struct Flat
{
    std::vector<Node> nodes;
    std::unordered_map<Task *, size_t> index;
    std::unordered_set<Task *> seenGraphs;

    Node &operator[](Task *t)
    {
        std::unordered_map<Task *, size_t>::iterator i = index.find(t);
        if (i != index.end()) {
            return nodes[i->second];
        }
        index.insert(std::make_pair(t, nodes.size()));
        nodes.push_back(Node());
        nodes.back().task = t;
        return nodes.back();
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
    if (!nodes.seenGraphs.insert(t).second) {
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

bool byContext(Task *a, Task *b)
{
    void *ca = a->getContext(), *cb = b->getContext();
    return ca != cb ? ca < cb : a < b;
}

}  // namespace

static void initAll(const std::vector<ptr<Task> > &roots, std::set<Task *> &initialized)
{
    for (;;) {
        Flat nodes;
        for (size_t i = 0; i < roots.size(); ++i) {
            flatten(roots[i].get(), nodes);
        }
        bool any = false;
        /* keep the tasks alive: an init may drop the last other reference to a sibling */
        std::vector<ptr<Task> > todo;
        for (std::vector<Node>::iterator n = nodes.nodes.begin(); n != nodes.nodes.end(); ++n) {
            if (initialized.insert(n->task).second) {
                todo.push_back(n->task);
            }
        }
        for (size_t i = 0; i < todo.size(); ++i) {
            todo[i]->init(initialized);
            any = true;
        }
        if (!any) {
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
    initAll(roots, initialized);

    for (;;) {
        /* flatten again every wave: finishing a task releases its inputs, restarting one re-acquires
         * them, both edit the graphs */
        Flat nodes;
        for (size_t i = 0; i < roots.size(); ++i) {
            flatten(roots[i].get(), nodes);
        }
        /* done tasks that completed before one of their inputs changed are stale */
        bool restarted = false;
        for (std::vector<Node>::iterator n = nodes.nodes.begin(); n != nodes.nodes.end(); ++n) {
            Task *t = n->task;
            if (!t->isDone()) continue;
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
            initAll(roots, initialized);
            continue;
        }

        std::vector<Task *> ready;
        size_t open = 0;
        for (std::vector<Node>::iterator n = nodes.nodes.begin(); n != nodes.nodes.end(); ++n) {
            if (n->task->isDone()) continue;
            ++open;
            bool ok = true;
            for (size_t d = 0; d < n->deps.size() && ok; ++d) {
                ok = n->deps[d]->isDone();
            }
            if (ok) ready.push_back(n->task);
        }
        if (open == 0) {
            break;
        }
        if (ready.empty()) {
            if (Logger::ERROR_LOGGER != NULL) {
                Logger::ERROR_LOGGER->logf("SCHEDULER", "%d tasks left but none can run (dependency cycle)", (int) open);
            }
            throw std::logic_error("BatchScheduler: dependency cycle");
        }
        std::sort(ready.begin(), ready.end(), byContext);

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
    for (size_t i = 0; i < roots.size(); ++i) {
        if (roots[i]->isTaskGraph()) {
            /* a graph has no run(); record its completion */
            roots[i]->Task::setIsDone(true, frame, Task::DATA_NEEDED);
        }
    }
}

}  // namespace ork
