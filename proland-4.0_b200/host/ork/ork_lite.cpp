/*
 * ork_lite.cpp -- Logger, Task and TaskGraph of ork_lite.h.
 */
#include "ork/ork_lite.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>

namespace ork
{

static Logger g_info("INFO"), g_warning("WARNING"), g_error("ERROR");
Logger *Logger::DEBUG_LOGGER = NULL;
Logger *Logger::INFO_LOGGER = &g_info;
Logger *Logger::WARNING_LOGGER = &g_warning;
Logger *Logger::ERROR_LOGGER = &g_error;

void Logger::log(const std::string &topic, const std::string &msg)
{
    ++lines;
    last = topic + ": " + msg;
    if (echo) {
        fprintf(stderr, "%s [%s] %s\n", level, topic.c_str(), msg.c_str());
    }
}

void Logger::logf(const char *topic, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    log(topic, buf);
}

Task::Task(const char *type, bool gpuTask, unsigned int deadline) :
    Object(type), schedViewStamp(0), schedSlot(0), schedMemoStamp(0), schedMemo(false), done(false), completionDate(0), changeDate(0),
    lastReason(DATA_NEEDED), gpuTask(gpuTask), deadline(deadline)
{
}

Task::~Task()
{
}

void Task::setIsDone(bool d, unsigned int t, reason r)
{
    done = d;
    if (d) {
        completionDate = t;
        if (lastReason == DATA_CHANGED) {
            changeDate = t;
        }
        lastReason = DATA_NEEDED;
    } else if (r == DATA_CHANGED || lastReason != DATA_CHANGED) {
        lastReason = r;
    }
}

std::atomic<unsigned long long> TaskGraph::edits(0);

TaskGraph::TaskGraph() : Task("TaskGraph", false, 0), index(NULL)
{
}

TaskGraph::~TaskGraph()
{
    delete index;
}

bool TaskGraph::isDone()
{
    for (TaskSet::const_iterator i = tasks.begin(); i != tasks.end(); ++i) {
        if (!(*i)->isDone()) {
            return false;
        }
    }
    return true;
}

void TaskGraph::setIsDone(bool d, unsigned int t, reason r)
{
    Task::setIsDone(d, t, r);
    if (!d) {
        /* the result of a graph is what its last tasks make: those must run again */
        TaskIterator i = getLastTasks();
        while (i.hasNext()) {
            i.next()->setIsDone(false, t, r);
        }
    }
}

unsigned int TaskGraph::getCompletionDate()
{
    unsigned int d = 0;
    for (TaskSet::const_iterator i = tasks.begin(); i != tasks.end(); ++i) {
        d = std::max(d, (*i)->getCompletionDate());
    }
    return d;
}

unsigned int TaskGraph::getChangeDate()
{
    /* what a graph produces is what its last tasks produce */
    unsigned int d = 0;
    TaskIterator i = getLastTasks();
    while (i.hasNext()) {
        d = std::max(d, i.next()->getChangeDate());
    }
    return d;
}

void TaskGraph::init(std::set<Task *> &initialized)
{
    initialized.insert(this);
    /* a task's init may edit this graph (CreateTile::start rebuilds its dependencies and adds the
     * tasks of the tiles it acquires): iterate snapshots until no task is left uninitialised.
     * Sub-graphs are always walked: they may hold tasks that were restarted since the last pass. */
    bool again = true;
    while (again) {
        again = false;
        std::vector<ptr<Task> > snapshot(tasks.begin(), tasks.end());
        for (size_t i = 0; i < snapshot.size(); ++i) {
            Task *t = snapshot[i].get();
            if (t->isTaskGraph()) {
                t->init(initialized);
            } else if (initialized.insert(t).second) {
                t->init(initialized);
                again = true;
            }
        }
    }
}

bool TaskGraph::contains(Task *t) const
{
    if (index != NULL) {
        return index->find(t) != index->end();
    }
    for (size_t i = 0; i < tasks.size(); ++i) {
        if (tasks[i].get() == t) return true;
    }
    return false;
}

TaskGraph::Needs *TaskGraph::needsOf(Task *src)
{
    for (size_t i = 0; i < dependencies.size(); ++i) {
        if (dependencies[i].src == src) return &dependencies[i];
    }
    return NULL;
}

const TaskGraph::Needs *TaskGraph::needsOf(Task *src) const
{
    return const_cast<TaskGraph *>(this)->needsOf(src);
}

TaskGraph::NeededBy *TaskGraph::neededBy(Task *dst)
{
    for (size_t i = 0; i < inverse.size(); ++i) {
        if (inverse[i].dst == dst) return &inverse[i];
    }
    return NULL;
}

const TaskGraph::NeededBy *TaskGraph::neededBy(Task *dst) const
{
    return const_cast<TaskGraph *>(this)->neededBy(dst);
}

void TaskGraph::addTask(ptr<Task> t)
{
    if (contains(t.get())) {
        return;
    }
    ++edits;
    tasks.push_back(t);
    if (index != NULL) {
        index->insert(t.get());
    } else if (tasks.size() > (size_t) kIndexAbove) {
        index = new std::unordered_set<Task *>();
        index->reserve(4 * tasks.size());
        for (size_t i = 0; i < tasks.size(); ++i) index->insert(tasks[i].get());
    }
}

void TaskGraph::removeTask(ptr<Task> t)
{
    ++edits;
    TaskSet gone;
    removeAndGetDependencies(t, gone);
    if (NeededBy *nb = neededBy(t.get())) {
        const std::vector<Task *> users = nb->src;
        for (size_t u = 0; u < users.size(); ++u) {
            if (Needs *n = needsOf(users[u])) {
                for (size_t k = 0; k < n->dst.size(); ++k) {
                    if (n->dst[k].get() == t.get()) {
                        n->dst.erase(n->dst.begin() + k);
                        break;
                    }
                }
            }
        }
        nb = neededBy(t.get());
        inverse.erase(inverse.begin() + (nb - &inverse[0]));
    }
    for (size_t i = 0; i < tasks.size(); ++i) {
        if (tasks[i].get() == t.get()) {
            tasks.erase(tasks.begin() + i);
            break;
        }
    }
    if (index != NULL) index->erase(t.get());
}

void TaskGraph::addDependency(ptr<Task> src, ptr<Task> dst)
{
    ++edits;
    Needs *n = needsOf(src.get());
    if (n == NULL) {
        dependencies.push_back(Needs());
        n = &dependencies.back();
        n->src = src.get();
    }
    bool have = false;
    for (size_t k = 0; k < n->dst.size() && !have; ++k) have = n->dst[k].get() == dst.get();
    if (!have) n->dst.push_back(dst);
    NeededBy *nb = neededBy(dst.get());
    if (nb == NULL) {
        inverse.push_back(NeededBy());
        nb = &inverse.back();
        nb->dst = dst.get();
    }
    if (std::find(nb->src.begin(), nb->src.end(), src.get()) == nb->src.end()) nb->src.push_back(src.get());
}

void TaskGraph::removeDependency(ptr<Task> src, ptr<Task> dst)
{
    ++edits;
    if (Needs *n = needsOf(src.get())) {
        for (size_t k = 0; k < n->dst.size(); ++k) {
            if (n->dst[k].get() == dst.get()) {
                n->dst.erase(n->dst.begin() + k);
                break;
            }
        }
        if (n->dst.empty()) dependencies.erase(dependencies.begin() + (n - &dependencies[0]));
    }
    if (NeededBy *nb = neededBy(dst.get())) {
        std::vector<Task *>::iterator i = std::find(nb->src.begin(), nb->src.end(), src.get());
        if (i != nb->src.end()) nb->src.erase(i);
        if (nb->src.empty()) inverse.erase(inverse.begin() + (nb - &inverse[0]));
    }
}

void TaskGraph::removeAndGetDependencies(ptr<Task> src, TaskSet &deletedDependencies)
{
    const Needs *n = needsOf(src.get());
    if (n == NULL) {
        return;
    }
    const TaskSet deps = n->dst;
    for (size_t i = 0; i < deps.size(); ++i) {
        bool have = false;
        for (size_t k = 0; k < deletedDependencies.size() && !have; ++k) have = deletedDependencies[k].get() == deps[i].get();
        if (!have) deletedDependencies.push_back(deps[i]);
        removeDependency(src, deps[i]);
    }
}

void TaskGraph::clearDependencies()
{
    ++edits;
    dependencies.clear();
    inverse.clear();
}

void TaskGraph::cleanup()
{
    ++edits;
    dependencies.clear();
    inverse.clear();
    tasks.clear();
    delete index;
    index = NULL;
}

TaskGraph::TaskIterator TaskGraph::getAllTasks() const
{
    return TaskIterator(tasks);
}

TaskGraph::TaskIterator TaskGraph::getFirstTasks() const
{
    std::vector<ptr<Task> > v;
    for (size_t i = 0; i < tasks.size(); ++i) {
        const Needs *n = needsOf(tasks[i].get());
        if (n == NULL || n->dst.empty()) v.push_back(tasks[i]);
    }
    return TaskIterator(v);
}

TaskGraph::TaskIterator TaskGraph::getLastTasks() const
{
    std::vector<ptr<Task> > v;
    for (size_t i = 0; i < tasks.size(); ++i) {
        const NeededBy *nb = neededBy(tasks[i].get());
        if (nb == NULL || nb->src.empty()) v.push_back(tasks[i]);
    }
    return TaskIterator(v);
}

TaskGraph::TaskIterator TaskGraph::getDependencies(ptr<Task> t) const
{
    const Needs *n = needsOf(t.get());
    return n == NULL ? TaskIterator() : TaskIterator(n->dst);
}

TaskGraph::TaskIterator TaskGraph::getInverseDependencies(ptr<Task> t) const
{
    std::vector<ptr<Task> > v;
    if (const NeededBy *nb = neededBy(t.get())) {
        for (size_t i = 0; i < nb->src.size(); ++i) v.push_back(ptr<Task>(nb->src[i]));
    }
    return TaskIterator(v);
}

const TaskGraph::TaskSet *TaskGraph::dependenciesOf(Task *t) const
{
    const Needs *n = needsOf(t);
    return n == NULL ? NULL : &n->dst;
}

}  // namespace ork
