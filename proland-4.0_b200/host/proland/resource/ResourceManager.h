/*
 * ResourceManager -- loads the tile-production resources of a Proland archive.
 *
 * The reference declares its producers as XML resources resolved by name through
 * Ork's ResourceManager (ElevationProducer.cpp:457-577, NormalProducer.cpp:333-383,
 * ResidualProducer.cpp:386-436, TileCache.cpp:438-468, GPUTileStorage.cpp:263-294,
 * CPUTileStorage.cpp:36-63).  This class keeps that surface for the hot path: the
 * same element names, the same attribute whitelists (an unknown attribute is an
 * error, like Resource::checkParameters), the same defaults, lazy creation by
 * name, and the `name suffix 1..6 -> cube face` rule.
 *
 *   <multithreadScheduler name nthreads fps prefetchRate prefetchQueue/>
 *   <tileCache name scheduler [storage]> <gpuTileStorage .../> | <cpuFloatTileStorage .../> </tileCache>
 *   <gpuTileStorage [name] tileSize nTiles internalformat format type min mag .../>
 *   <cpuFloatTileStorage [name] tileSize channels capacity/>
 *   <residualProducer name cache [file] [delta] [scale]> nested <residualProducer>s </residualProducer>
 *   <elevationProducer name cache [residuals] [face] [upsampleProg] [blendProg] [gridSize] [noise] [flip]/>
 *   <normalProducer name cache elevations [normalProg] [gridSize] [deform]/>
 *
 * Elements of other plugins in the same archive (ortho, terrainNode, ...) are left alone.
 */
#ifndef PROLAND_B200_RESOURCE_MANAGER_H
#define PROLAND_B200_RESOURCE_MANAGER_H

#include <map>
#include <string>
#include <vector>

#include "ork/ork_lite.h"
#include "proland/resource/XmlLite.h"

using namespace ork;

namespace proland
{

class ResourceManager : public Object
{
public:
    /* dataDir: where `file=` attributes are looked up (Ork's resource path) */
    ResourceManager(const std::string &archiveXml, const std::string &dataDir = ".", int device = -1);
    virtual ~ResourceManager();

    /* the resource called `name`, created on first use; throws std::runtime_error when the archive has
     * no such resource or its description is invalid */
    ptr<Object> loadResource(const std::string &name);
    bool hasResource(const std::string &name) const;
    std::vector<std::string> getResourceNames() const;
    /* drops every resource (in reverse creation order) */
    void close();

    /* upsampleProg -> shader variant; "name;" as the reference spells program lists */
    static bool upsampleVariant(const std::string &prog, bool *slopeNoise, bool *noClamp);

private:
    XmlElement archive;
    std::string dataDir;
    int device;
    std::map<std::string, const XmlElement *> descriptors;
    std::map<std::string, ptr<Object> > resources;
    std::vector<std::string> order;
    std::vector<std::string> loading;   /* cycle detection */

    ptr<Object> create(const std::string &name, const XmlElement *e);
    ptr<Object> createStorage(const XmlElement *e);
    std::string findFile(const std::string &file) const;
};

}  // namespace proland

#endif
