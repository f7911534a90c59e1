#include "proland/resource/ResourceManager.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <sys/stat.h>

#include "ork/BatchScheduler.h"
#include "proland/dem/ElevationProducer.h"
#include "proland/dem/NormalProducer.h"
#include "proland/ortho/OrthoCPUProducer.h"
#include "proland/ortho/OrthoProducer.h"
#include "proland/dem/ResidualProducer.h"
#include "proland/producer/CPUTileStorage.h"
#include "proland/producer/GPUTileStorage.h"
#include "proland/producer/TileCache.h"

namespace proland
{

namespace
{

void fail(const XmlElement *e, const std::string &msg)
{
    /* no iostreams in this library: it is loaded into processes (Python + numpy) whose extension
     * modules carry their own libstdc++ locale objects */
    const std::string o = "<" + e->name + "> (line " + std::to_string(e->line) + "): " + msg;
    if (Logger::ERROR_LOGGER != NULL) {
        Logger::ERROR_LOGGER->log("RESOURCE", o);
    }
    throw std::runtime_error(o);
}

/* Resource::checkParameters: every attribute must be in the comma-terminated whitelist */
void checkParameters(const XmlElement *e, const char *params)
{
    for (size_t i = 0; i < e->attributes.size(); ++i) {
        const std::string key = e->attributes[i].first + ",";
        const char *hit = strstr(params, key.c_str());
        bool ok = false;
        while (hit != NULL && !ok) {
            ok = hit == params || hit[-1] == ',';
            if (!ok) hit = strstr(hit + 1, key.c_str());
        }
        if (!ok) {
            fail(e, "unsupported '" + e->attributes[i].first + "' attribute");
        }
    }
}

std::string getParameter(const XmlElement *e, const char *name)
{
    const char *v = e->Attribute(name);
    if (v == NULL) {
        fail(e, std::string("missing '") + name + "' attribute");
    }
    return v;
}

void getIntParameter(const XmlElement *e, const char *name, int *out)
{
    const std::string v = getParameter(e, name);
    char *end = NULL;
    const long r = strtol(v.c_str(), &end, 10);
    if (end == v.c_str() || *end != 0) {
        fail(e, std::string("invalid integer '") + name + "' attribute");
    }
    *out = (int) r;
}

void getFloatParameter(const XmlElement *e, const char *name, float *out)
{
    const std::string v = getParameter(e, name);
    char *end = NULL;
    const float r = strtof(v.c_str(), &end);
    if (end == v.c_str() || *end != 0) {
        fail(e, std::string("invalid float '") + name + "' attribute");
    }
    *out = r;
}

bool isStorage(const std::string &n)
{
    return n == "gpuTileStorage" || n == "cpuFloatTileStorage" || n == "cpuByteTileStorage";
}

bool isKnown(const std::string &n)
{
    return isStorage(n) || n == "multithreadScheduler" || n == "tileCache" || n == "residualProducer" ||
           n == "elevationProducer" || n == "normalProducer" || n == "orthoProducer" || n == "orthoCpuProducer";
}

}  // namespace

bool ResourceManager::upsampleVariant(const std::string &prog, bool *slopeNoise, bool *noClamp)
{
    std::string p = prog;
    while (!p.empty() && p[p.size() - 1] == ';') p.erase(p.size() - 1);
    /* demo/shaders/elevation: upsampleShader.xml, upsampleShader-noClamp.xml (variant D);
     * the examples' plain-noise shaders are selected as upsampleShader-plain[-noClamp] */
    *slopeNoise = p.find("-plain") == std::string::npos;
    *noClamp = p.find("-noClamp") != std::string::npos;
    std::string base = p;
    const size_t dash = base.find('-');
    if (dash != std::string::npos) base.erase(dash);
    return base == "upsampleShader";
}

ResourceManager::ResourceManager(const std::string &archiveXml, const std::string &dataDir, int device) :
    Object("ResourceManager"), dataDir(dataDir), device(device)
{
    archive = parseXml(archiveXml);
    /* an archive is a flat list of named resources; a single resource may also be the root */
    std::vector<const XmlElement *> top;
    if (archive.name == "archive") {
        for (size_t i = 0; i < archive.children.size(); ++i) top.push_back(&archive.children[i]);
    } else {
        top.push_back(&archive);
    }
    for (size_t i = 0; i < top.size(); ++i) {
        const char *n = top[i]->Attribute("name");
        if (n != NULL && isKnown(top[i]->name)) {
            if (descriptors.find(n) != descriptors.end()) {
                fail(top[i], std::string("duplicate resource name '") + n + "'");
            }
            descriptors[n] = top[i];
        }
    }
}

ResourceManager::~ResourceManager()
{
    close();
}

void ResourceManager::close()
{
    /* queued prefetch tasks pin tiles: drop them first */
    for (std::map<std::string, ptr<Object> >::iterator i = resources.begin(); i != resources.end(); ++i) {
        BatchScheduler *s = dynamic_cast<BatchScheduler *>(i->second.get());
        if (s != NULL) {
            s->clear();
        }
    }
    /* users before what they use */
    while (!order.empty()) {
        resources.erase(order.back());
        order.pop_back();
    }
    resources.clear();
}

bool ResourceManager::hasResource(const std::string &name) const
{
    return descriptors.find(name) != descriptors.end();
}

std::vector<std::string> ResourceManager::getResourceNames() const
{
    std::vector<std::string> r;
    for (std::map<std::string, const XmlElement *>::const_iterator i = descriptors.begin(); i != descriptors.end(); ++i) {
        r.push_back(i->first);
    }
    return r;
}

std::string ResourceManager::findFile(const std::string &file) const
{
    struct stat st;
    if (!file.empty() && file[0] == '/' && stat(file.c_str(), &st) == 0) {
        return file;
    }
    const std::string p = dataDir + "/" + file;
    if (stat(p.c_str(), &st) == 0) {
        return p;
    }
    return file;   /* the producer reports "Cannot open file" like the reference */
}

ptr<Object> ResourceManager::loadResource(const std::string &name)
{
    std::map<std::string, ptr<Object> >::iterator r = resources.find(name);
    if (r != resources.end()) {
        return r->second;
    }
    std::map<std::string, const XmlElement *>::iterator d = descriptors.find(name);
    if (d == descriptors.end()) {
        if (Logger::ERROR_LOGGER != NULL) {
            Logger::ERROR_LOGGER->log("RESOURCE", "Missing or invalid resource '" + name + "'");
        }
        throw std::runtime_error("Missing or invalid resource '" + name + "'");
    }
    if (std::find(loading.begin(), loading.end(), name) != loading.end()) {
        fail(d->second, "resource '" + name + "' depends on itself");
    }
    loading.push_back(name);
    ptr<Object> o;
    try {
        o = create(name, d->second);
    } catch (...) {
        loading.pop_back();
        throw;
    }
    loading.pop_back();
    resources[name] = o;
    order.push_back(name);
    return o;
}

ptr<Object> ResourceManager::createStorage(const XmlElement *e)
{
    ptr<DeviceContext> context = DeviceContext::get(device);
    if (e->name == "gpuTileStorage") {
        checkParameters(e, "name,tileSize,nTiles,tileMap,internalformat,format,type,min,mag,minLod,maxLod,minLevel,maxLevel,swizzle,anisotropy,");
        int tileSize, nTiles;
        getIntParameter(e, "tileSize", &tileSize);
        getIntParameter(e, "nTiles", &nTiles);
        TextureInternalFormat tf;
        if (!GPUTileStorage::parseInternalFormat(getParameter(e, "internalformat"), &tf)) {
            fail(e, "internalformat must be RGB32F, RGBA32F (elevations), RG8, RGBA8 (normals) or RGB8, RGBA8 (ortho) on this path");
        }
        TextureFilter minf = NEAREST, magf = NEAREST;
        if (e->Attribute("min") != NULL && !GPUTileStorage::parseFilter(e->Attribute("min"), &minf)) {
            fail(e, "min filter must be NEAREST, LINEAR or one of their MIPMAP variants");
        }
        if (e->Attribute("mag") != NULL && !GPUTileStorage::parseFilter(e->Attribute("mag"), &magf)) {
            fail(e, "mag filter must be NEAREST or LINEAR");
        }
        return new GPUTileStorage(tileSize, nTiles, tf, minf, magf, context);
    }
    checkParameters(e, "name,tileSize,channels,capacity,");
    int tileSize, channels, capacity;
    getIntParameter(e, "tileSize", &tileSize);
    getIntParameter(e, "channels", &channels);
    getIntParameter(e, "capacity", &capacity);
    if (e->name == "cpuByteTileStorage") {
        return new CPUTileStorage<unsigned char>(tileSize, channels, capacity, context);
    }
    return new CPUTileStorage<float>(tileSize, channels, capacity, context);
}

ptr<Object> ResourceManager::create(const std::string &name, const XmlElement *e)
{
    if (e->name == "multithreadScheduler") {
        checkParameters(e, "name,nthreads,fps,prefetchRate,prefetchQueue,");
        int rate = 0, queue = 0;
        if (e->Attribute("prefetchRate") != NULL) getIntParameter(e, "prefetchRate", &rate);
        if (e->Attribute("prefetchQueue") != NULL) getIntParameter(e, "prefetchQueue", &queue);
        return new BatchScheduler(rate, queue);
    }
    if (isStorage(e->name)) {
        return createStorage(e);
    }
    if (e->name == "tileCache") {
        checkParameters(e, "name,storage,scheduler,");
        ptr<TileStorage> storage;
        if (e->Attribute("storage") != NULL) {
            storage = loadResource(getParameter(e, "storage")).cast<TileStorage>();
        } else {
            if (e->children.empty()) {
                fail(e, "Missing storage attribute or subelement");
            }
            if (!isStorage(e->children[0].name)) {
                fail(&e->children[0], "not a tile storage of the tile-production path");
            }
            storage = createStorage(&e->children[0]).cast<TileStorage>();
        }
        if (storage == NULL) fail(e, "storage is not a TileStorage");
        ptr<Scheduler> scheduler = loadResource(getParameter(e, "scheduler")).cast<Scheduler>();
        if (scheduler == NULL) fail(e, "scheduler is not a Scheduler");
        return new TileCache(storage, name, scheduler);
    }
    if (e->name == "residualProducer") {
        checkParameters(e, "name,cache,file,delta,scale,");
        ptr<TileCache> cache = loadResource(getParameter(e, "cache")).cast<TileCache>();
        if (cache == NULL) fail(e, "cache is not a TileCache");
        std::string file;
        int deltaLevel = 0;
        float zscale = 1.0f;
        if (e->Attribute("file") != NULL) file = findFile(getParameter(e, "file"));
        if (e->Attribute("scale") != NULL) getFloatParameter(e, "scale", &zscale);
        if (e->Attribute("delta") != NULL) getIntParameter(e, "delta", &deltaLevel);
        ptr<ResidualProducer> p = new ResidualProducer(cache, file.c_str(), deltaLevel, zscale);
        for (size_t i = 0; i < e->children.size(); ++i) {
            const XmlElement *f = &e->children[i];
            if (strncmp(f->name.c_str(), "residualProducer", 16) == 0) {
                const char *childName = f->Attribute("name");
                p->addProducer(create(childName ? childName : "", f).cast<ResidualProducer>());
            } else {
                fail(f, "Invalid subelement");
            }
        }
        return p;
    }
    if (e->name == "elevationProducer") {
        checkParameters(e, "name,cache,residuals,face,upsampleProg,blendProg,gridSize,noise,flip,");
        ptr<TileCache> cache = loadResource(getParameter(e, "cache")).cast<TileCache>();
        if (cache == NULL) fail(e, "cache is not a TileCache");
        ptr<TileProducer> residuals;
        if (e->Attribute("residuals") != NULL) {
            residuals = loadResource(getParameter(e, "residuals")).cast<TileProducer>();
            if (residuals == NULL) fail(e, "residuals is not a TileProducer");
        }
        std::string upsample = "upsampleShader;";
        if (e->Attribute("upsampleProg") != NULL) upsample = getParameter(e, "upsampleProg");
        bool slopeNoise, noClamp;
        if (!upsampleVariant(upsample, &slopeNoise, &noClamp)) {
            fail(e, "unknown upsampleProg '" + upsample + "' (upsampleShader[-plain][-noClamp];)");
        }
        int gridSize = 24;
        if (e->Attribute("gridSize") != NULL) getIntParameter(e, "gridSize", &gridSize);
        std::vector<float> noiseAmp;
        if (e->Attribute("noise") != NULL) {
            /* comma separated, sscanf("%f") per item like the reference (ElevationProducer.cpp:481-492) */
            const std::string noiseAmps = std::string(e->Attribute("noise")) + ",";
            std::string::size_type start = 0, index;
            while ((index = noiseAmps.find(',', start)) != std::string::npos) {
                float value = 0.0f;
                sscanf(noiseAmps.substr(start, index - start).c_str(), "%f", &value);
                noiseAmp.push_back(value);
                start = index + 1;
            }
        }
        const bool flip = e->Attribute("flip") != NULL && strcmp(e->Attribute("flip"), "true") == 0;
        int face = 0;
        if (e->Attribute("face") != NULL) {
            getIntParameter(e, "face", &face);
        } else if (!name.empty() && name[name.size() - 1] >= '1' && name[name.size() - 1] <= '6') {
            face = name[name.size() - 1] - '0';
        }
        for (size_t i = 0; i < e->children.size(); ++i) {
            /* layers need the graph plugin: out of scope */
            if (Logger::WARNING_LOGGER != NULL) {
                Logger::WARNING_LOGGER->log("RESOURCE", "Unknown scene node element '" + e->children[i].name + "'");
            }
        }
        return new ElevationProducer(cache, residuals, gridSize, noiseAmp, flip, face,
                                     ElevationProducer::Variant(slopeNoise, noClamp));
    }
    if (e->name == "normalProducer") {
        checkParameters(e, "name,cache,elevations,normalProg,gridSize,deform,");
        ptr<TileCache> cache = loadResource(getParameter(e, "cache")).cast<TileCache>();
        if (cache == NULL) fail(e, "cache is not a TileCache");
        ptr<TileProducer> elevations = loadResource(getParameter(e, "elevations")).cast<TileProducer>();
        if (elevations == NULL) fail(e, "elevations is not a TileProducer");
        int gridSize = 24;
        if (e->Attribute("gridSize") != NULL) getIntParameter(e, "gridSize", &gridSize);
        const bool deform = e->Attribute("deform") != NULL && strcmp(e->Attribute("deform"), "sphere") == 0;
        return new NormalProducer(cache, elevations, gridSize, deform);
    }
    if (e->name == "orthoCpuProducer") {
        /* OrthoCPUProducer.cpp:248-266 */
        checkParameters(e, "name,cache,file,");
        ptr<TileCache> cache = loadResource(getParameter(e, "cache")).cast<TileCache>();
        if (cache == NULL) fail(e, "cache is not a TileCache");
        std::string file;
        if (e->Attribute("file") != NULL) file = findFile(getParameter(e, "file"));
        return new OrthoCPUProducer(cache, file.c_str());
    }
    if (e->name == "orthoProducer") {
        /* OrthoProducer.cpp:440-512 */
        checkParameters(e, "name,cache,residuals,face,upsampleProg,rnoise,cnoise,noise,hsv,scale,maxLevel,");
        ptr<TileCache> cache = loadResource(getParameter(e, "cache")).cast<TileCache>();
        if (cache == NULL) fail(e, "cache is not a TileCache");
        ptr<TileProducer> residuals;
        if (e->Attribute("residuals") != NULL) {
            residuals = loadResource(getParameter(e, "residuals")).cast<TileProducer>();
            if (residuals == NULL) fail(e, "residuals is not a TileProducer");
        }
        if (e->Attribute("upsampleProg") != NULL) {
            std::string prog = getParameter(e, "upsampleProg");
            while (!prog.empty() && prog[prog.size() - 1] == ';') prog.erase(prog.size() - 1);
            if (prog != "upsampleOrthoShader") fail(e, "unknown upsampleProg '" + prog + "' (upsampleOrthoShader;)");
        }
        float rootNoiseColor[4] = { 0.5f, 0.5f, 0.5f, 0.5f };
        float noiseColor[4] = { 1.0f, 1.0f, 1.0f, 1.0f };
        const char *colorAttr[2] = { "rnoise", "cnoise" };
        float *colors[2] = { rootNoiseColor, noiseColor };
        for (int k = 0; k < 2; ++k) {
            if (e->Attribute(colorAttr[k]) == NULL) continue;
            /* four items always: a missing one is atof("") = 0 (OrthoProducer.cpp:462-481) */
            const std::string c = std::string(e->Attribute(colorAttr[k])) + ",";
            std::string::size_type start = 0;
            for (int i = 0; i < 4; ++i) {
                const std::string::size_type index = c.find(',', start);
                const std::string item = index == std::string::npos ? std::string() : c.substr(start, index - start);
                colors[k][i] = (float) atof(item.c_str()) / 255;
                start = index == std::string::npos ? c.size() : index + 1;
            }
        }
        std::vector<float> noiseAmp;
        if (e->Attribute("noise") != NULL) {
            const std::string noiseAmps = std::string(e->Attribute("noise")) + ",";
            std::string::size_type start = 0, index;
            while ((index = noiseAmps.find(',', start)) != std::string::npos) {
                float value = 0.0f;
                sscanf(noiseAmps.substr(start, index - start).c_str(), "%f", &value);
                noiseAmp.push_back(value);
                start = index + 1;
            }
        }
        const bool hsv = e->Attribute("hsv") != NULL && strcmp(e->Attribute("hsv"), "true") == 0;
        float scale = 2.0f;
        int maxLevel = -1;
        if (e->Attribute("scale") != NULL) getFloatParameter(e, "scale", &scale);
        if (e->Attribute("maxLevel") != NULL) getIntParameter(e, "maxLevel", &maxLevel);
        int face = 1;
        if (e->Attribute("face") != NULL) {
            getIntParameter(e, "face", &face);
        } else if (!name.empty() && name[name.size() - 1] >= '1' && name[name.size() - 1] <= '6') {
            face = name[name.size() - 1] - '0';
        }
        for (size_t i = 0; i < e->children.size(); ++i) {
            if (Logger::WARNING_LOGGER != NULL) {
                Logger::WARNING_LOGGER->log("RESOURCE", "Unknown scene node element '" + e->children[i].name + "'");
            }
        }
        return new OrthoProducer(cache, residuals, rootNoiseColor, noiseColor, noiseAmp, hsv, scale, maxLevel, face);
    }
    fail(e, "not a resource of the tile-production path");
    return NULL;
}

}  // namespace proland
