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

unsigned long long TaskGraph::edits = 0;

TaskGraph::TaskGraph() : Task("TaskGraph", false, 0)
{
}

TaskGraph::~TaskGraph()
{
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

void TaskGraph::addTask(ptr<Task> t)
{
    ++edits;
    tasks.insert(t);
}

void TaskGraph::removeTask(ptr<Task> t)
{
    ++edits;
    TaskSet gone;
    removeAndGetDependencies(t, gone);
    std::map<Task *, std::set<Task *> >::iterator inv = inverse.find(t.get());
    if (inv != inverse.end()) {
        std::set<Task *> users = inv->second;
        for (std::set<Task *>::iterator u = users.begin(); u != users.end(); ++u) {
            dependencies[*u].erase(t);
        }
        inverse.erase(t.get());
    }
    tasks.erase(t);
}

void TaskGraph::addDependency(ptr<Task> src, ptr<Task> dst)
{
    ++edits;
    dependencies[src.get()].insert(dst);
    inverse[dst.get()].insert(src.get());
}

void TaskGraph::removeDependency(ptr<Task> src, ptr<Task> dst)
{
    ++edits;
    std::map<Task *, TaskSet>::iterator d = dependencies.find(src.get());
    if (d != dependencies.end()) {
        d->second.erase(dst);
        if (d->second.empty()) dependencies.erase(d);
    }
    std::map<Task *, std::set<Task *> >::iterator i = inverse.find(dst.get());
    if (i != inverse.end()) {
        i->second.erase(src.get());
        if (i->second.empty()) inverse.erase(i);
    }
}

void TaskGraph::removeAndGetDependencies(ptr<Task> src, TaskSet &deletedDependencies)
{
    std::map<Task *, TaskSet>::iterator d = dependencies.find(src.get());
    if (d == dependencies.end()) {
        return;
    }
    TaskSet deps = d->second;
    for (TaskSet::iterator i = deps.begin(); i != deps.end(); ++i) {
        deletedDependencies.insert(*i);
        removeDependency(src, *i);
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
}

TaskGraph::TaskIterator TaskGraph::getAllTasks() const
{
    return TaskIterator(std::vector<ptr<Task> >(tasks.begin(), tasks.end()));
}

TaskGraph::TaskIterator TaskGraph::getFirstTasks() const
{
    std::vector<ptr<Task> > v;
    for (TaskSet::const_iterator i = tasks.begin(); i != tasks.end(); ++i) {
        std::map<Task *, TaskSet>::const_iterator d = dependencies.find(i->get());
        if (d == dependencies.end() || d->second.empty()) v.push_back(*i);
    }
    return TaskIterator(v);
}

TaskGraph::TaskIterator TaskGraph::getLastTasks() const
{
    std::vector<ptr<Task> > v;
    for (TaskSet::const_iterator i = tasks.begin(); i != tasks.end(); ++i) {
        std::map<Task *, std::set<Task *> >::const_iterator d = inverse.find(i->get());
        if (d == inverse.end() || d->second.empty()) v.push_back(*i);
    }
    return TaskIterator(v);
}

TaskGraph::TaskIterator TaskGraph::getDependencies(ptr<Task> t) const
{
    std::map<Task *, TaskSet>::const_iterator d = dependencies.find(t.get());
    if (d == dependencies.end()) return TaskIterator();
    return TaskIterator(std::vector<ptr<Task> >(d->second.begin(), d->second.end()));
}

TaskGraph::TaskIterator TaskGraph::getInverseDependencies(ptr<Task> t) const
{
    std::vector<ptr<Task> > v;
    std::map<Task *, std::set<Task *> >::const_iterator d = inverse.find(t.get());
    if (d != inverse.end()) {
        for (std::set<Task *>::const_iterator i = d->second.begin(); i != d->second.end(); ++i) v.push_back(ptr<Task>(*i));
    }
    return TaskIterator(v);
}

const TaskGraph::TaskSet *TaskGraph::dependenciesOf(Task *t) const
{
    std::map<Task *, TaskSet>::const_iterator d = dependencies.find(t);
    return d == dependencies.end() ? NULL : &d->second;
}

}  // namespace ork
