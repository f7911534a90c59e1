/* Host cost of the plugin path, measured WITHOUT a GPU: ResourceManager -> TileCache -> Elevation / NormalProducer ->
 * getTile -> BatchScheduler::run -> doCreateTile, linked against host_path_stub.c (every device entry point of the C ABI
 * succeeds and does nothing) and the real pl_hostmath.cu (cnoise, fp64 patch geometry).  A profiling tool for the build
 * container (gprof has nothing else to look at there); nothing in the product or the tests loads the stub.
 *
 *   H=proland-4.0_b200/host; cd tools/microbench
 *   nvcc -O2 -std=c++17 --fmad=false -Xcompiler -ffp-contract=off -I../../include -I../../proland-4.0_b200/csrc \
 *        -c ../../proland-4.0_b200/csrc/pl_hostmath.cu -o /tmp/hostmath.o && gcc -O2 -w -c host_path_stub.c -o /tmp/stub.o
 *   g++ -O2 -g [-pg] -std=c++17 -UNDEBUG -I../../$H -I../../include host_path.cpp $(ls ../../$H/ork/*.cpp ../../$H/proland/[a-z]*[!s]/*.cpp \
 *        ../../$H/proland/*/*/*.cpp) /tmp/hostmath.o /tmp/stub.o -o /tmp/host_path -pthread -L/usr/local/cuda/lib64 -lcudart_static -ldl -lrt
 *   /tmp/host_path 6 5        # all 4 096 normal tiles of level 6 (+ 5 461 elevation tiles), 5 repetitions
 *
 * Round 2, 8 shared cores of the build container: 95 K tiles/s before (std::map tile ids, the flattened view rebuilt
 * every wave through two hash tables, TaskGraph::isDone re-walking every ancestor chain), 270 K tiles/s after, 300 K with
 * TaskGraph on flat vectors.  Requesting level L + 1 quadrants in turn against a cache of 12 000 tiles (every request evicts)
 * showed TileProducer::removeCreateTile's linear search: 45 K -> 100 K tiles/s with a hash set. */
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ork/BatchScheduler.h"
#include "proland/producer/TileCache.h"
#include "proland/producer/TileProducer.h"
#include "proland/resource/ResourceManager.h"
using namespace proland;
using namespace ork;
static const char *XML = R"(<?xml version="1.0" ?>
<archive>
    <multithreadScheduler name="defaultScheduler" nthreads="3" fps="0"/>
    <tileCache name="groundElevations" scheduler="defaultScheduler">
        <gpuTileStorage tileSize="101" nTiles="30000" internalformat="RGB32F" format="RGB" type="FLOAT" min="LINEAR" mag="LINEAR"/>
    </tileCache>
    <elevationProducer name="groundElevations1" cache="groundElevations" noise="-3250,-1590,-1125,-795,-561,-397,-140,-100,15,8,5,2.5,1.5,1,0.5,0.25,0.1,0.05"/>
    <tileCache name="groundNormals" scheduler="defaultScheduler">
        <gpuTileStorage tileSize="97" nTiles="30000" internalformat="RG8" format="RG" type="FLOAT" min="LINEAR" mag="LINEAR"/>
    </tileCache>
    <normalProducer name="groundNormals1" cache="groundNormals" elevations="groundElevations1" deform="sphere"/>
</archive>)";
int main(int argc, char **argv)
{
    int L = argc > 1 ? atoi(argv[1]) : 6, reps = argc > 2 ? atoi(argv[2]) : 5;
    ptr<ResourceManager> m = new ResourceManager(XML, ".", 0);
    TileProducer *norm = dynamic_cast<TileProducer *>(m->loadResource("groundNormals1").get());
    TileProducer *elev = dynamic_cast<TileProducer *>(m->loadResource("groundElevations1").get());
    elev->setRootQuadSize(12720000.0f);
    norm->setRootQuadSize(12720000.0f);
    ptr<Scheduler> s = norm->getCache()->getScheduler();
    int n = 1 << L;
    double best = 1e9;
    for (int r = 0; r < reps; ++r) {
        auto t0 = std::chrono::steady_clock::now();
        std::vector<TileCache::Tile *> tiles;
        ptr<TaskGraph> g = new TaskGraph();
        for (int ty = 0; ty < n; ++ty)
            for (int tx = 0; tx < n; ++tx) {
                TileCache::Tile *t = norm->getTile(L, tx, ty, 0);
                if (!t) { printf("cache full\n"); return 1; }
                tiles.push_back(t);
                g->addTask(t->task);
            }
        auto t1 = std::chrono::steady_clock::now();
        s->run(g);
        auto t2 = std::chrono::steady_clock::now();
        for (size_t i = 0; i < tiles.size(); ++i) norm->putTile(tiles[i]);
        g = NULL;
        norm->invalidateTiles();
        elev->invalidateTiles();
        auto t3 = std::chrono::steady_clock::now();
        double a = std::chrono::duration<double>(t1 - t0).count(), b = std::chrono::duration<double>(t2 - t1).count(), c = std::chrono::duration<double>(t3 - t2).count();
        size_t made = 0; for (int l = 0; l <= L; ++l) made += (size_t) 1 << (2 * l);
        printf("rep %d: getTile %.1f ms, run %.1f ms, put+invalidate %.1f ms: %.0f normal tiles/s requested, %.0f tiles/s made (elev+norm %zu)\n", r, a * 1e3, b * 1e3, c * 1e3,
               tiles.size() / (a + b + c), (made + tiles.size()) / (a + b + c), made + tiles.size());
        if (a + b + c < best) best = a + b + c;
    }
    m->close();
    return 0;
}
