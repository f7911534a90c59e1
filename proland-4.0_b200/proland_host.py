"""ctypes view of libproland_host.so: the C++ host layer (namespace proland -- TileStorage,
TileCache, TileProducer, Elevation/Normal/ResidualProducer, ResourceManager, BatchScheduler)
that mirrors the reference's producer interface on top of the C ABI (libproland_b200.so).

    scene = Scene(open("fractalterrain.xml").read())
    normals = scene.producer("groundNormals"); normals.set_root_quad_size(100000.0)
    tiles = [normals.get_tile(3, tx, ty) for ...]      # TileCache::getTile -> task graphs
    scene.run(tiles)                                    # Scheduler::run: one launch per producer per wave
    rg8 = tiles[0].download()

There is no CPU fallback: without a CUDA device opening a scene with GPU storages raises HostError.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libproland_host.so")
_lib = None

EXPORTS = [
    "plh_last_error", "plh_open", "plh_close", "plh_shutdown", "plh_producer", "plh_cache", "plh_scheduler",
    "plh_producer_cache", "plh_cache_scheduler", "plh_set_root_quad_size", "plh_producer_info", "plh_producer_type",
    "plh_producer_task_type", "plh_has_tile", "plh_has_children", "plh_get_tile", "plh_find_tile", "plh_put_tile",
    "plh_prefetch_tile", "plh_invalidate_tiles", "plh_invalidate_tile", "plh_producer_counts", "plh_tile_done",
    "plh_tile_slot", "plh_tile_download", "plh_tile_minmax", "plh_run", "plh_cache_stats", "plh_scheduler_stats",
    "plh_residual_info", "plh_residual_tile_id", "plh_residual_tile_size", "plh_device_launches", "plh_device_sync",
    "plh_debug_log", "plh_debug_log_lines", "plh_quiet_errors", "plh_upsample_variant", "plh_test_scene",
    "plh_test_scene_close", "plh_test_producer", "plh_test_cache", "plh_test_scheduler", "plh_test_calls",
    "plh_test_begin_end", "plh_terrain_create", "plh_terrain_destroy", "plh_split_distance", "plh_terrain_update",
    "plh_terrain_quads", "plh_terrain_quads_z", "plh_sampler_z_create", "plh_sampler_z_counts", "plh_ground_height", "plh_sampler_create", "plh_sampler_destroy", "plh_sampler_tile_count", "plh_frame_update",
    "plh_preprocess_dem",
]


def preprocess_dem(src, min_tile_size, tile_size, max_level, dst_folder, residual_scale=1.0, spherical=False):
    """proland::preprocessDem / preprocessSphericalDem on a (h, w) float32 map -> dst_folder/DEM.dat or DEM1..6.dat"""
    import numpy as np
    src = np.ascontiguousarray(src, np.float32)
    if lib().plh_preprocess_dem(src.ctypes.data, src.shape[1], src.shape[0], min_tile_size, tile_size, max_level,
                                os.fsencode(dst_folder), residual_scale, 1 if spherical else 0) != 0:
        raise HostError(_err())


class HostError(RuntimeError):
    pass


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "libproland_host.so"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HostError("%s is missing: run `make -C %s` (there is no fallback path)" % (LIB_PATH, HERE))
        L = C.CDLL(LIB_PATH)
        vp, i, u = C.c_void_p, C.c_int, C.c_uint
        L.plh_last_error.restype = C.c_char_p
        for name in ("plh_open", "plh_producer", "plh_cache", "plh_scheduler", "plh_producer_cache", "plh_cache_scheduler",
                     "plh_get_tile", "plh_find_tile", "plh_test_scene", "plh_test_producer", "plh_test_cache",
                     "plh_test_scheduler", "plh_terrain_create", "plh_sampler_create", "plh_sampler_z_create"):
            getattr(L, name).restype = vp
        L.plh_producer_type.restype = C.c_char_p
        L.plh_producer_task_type.restype = C.c_char_p
        L.plh_debug_log_lines.restype = C.c_ulong
        L.plh_device_launches.restype = C.c_ulonglong
        L.plh_open.argtypes = [C.c_char_p, C.c_char_p, i]
        L.plh_close.argtypes = [vp]
        for name in ("plh_producer", "plh_cache", "plh_scheduler"):
            getattr(L, name).argtypes = [vp, C.c_char_p]
        for name in ("plh_producer_cache", "plh_cache_scheduler", "plh_producer_type", "plh_producer_task_type",
                     "plh_invalidate_tiles", "plh_tile_done", "plh_tile_slot", "plh_test_scene_close", "plh_test_producer",
                     "plh_test_cache", "plh_test_scheduler"):
            getattr(L, name).argtypes = [vp]
        L.plh_set_root_quad_size.argtypes = [vp, C.c_float]
        L.plh_preprocess_dem.argtypes = [vp, i, i, i, i, i, C.c_char_p, C.c_float, i]
        for name in ("plh_producer_info", "plh_producer_counts", "plh_cache_stats", "plh_scheduler_stats", "plh_residual_info",
                     "plh_test_begin_end"):
            getattr(L, name).argtypes = [vp, vp]
        for name in ("plh_has_tile", "plh_has_children", "plh_prefetch_tile", "plh_invalidate_tile", "plh_residual_tile_id"):
            getattr(L, name).argtypes = [vp, i, i, i]
        L.plh_get_tile.argtypes = [vp, i, i, i, u]
        L.plh_find_tile.argtypes = [vp, i, i, i, i, i]
        L.plh_put_tile.argtypes = [vp, vp]
        L.plh_tile_download.argtypes = [vp, vp, C.c_size_t]
        L.plh_tile_minmax.argtypes = [vp, vp, vp]
        L.plh_run.argtypes = [vp, vp, i]
        L.plh_residual_tile_size.argtypes = [vp, i]
        L.plh_device_launches.argtypes = [i]
        L.plh_device_sync.argtypes = [i]
        L.plh_debug_log.argtypes = [i, i]
        L.plh_quiet_errors.argtypes = [i]
        L.plh_upsample_variant.argtypes = [C.c_char_p, vp]
        L.plh_test_scene.argtypes = [i, i, i, i, i]
        L.plh_terrain_create.argtypes = [C.c_float, C.c_float, C.c_float, C.c_float, i]
        L.plh_terrain_destroy.argtypes = [vp]
        L.plh_split_distance.argtypes = [C.c_float, C.c_float, C.c_float]
        L.plh_split_distance.restype = C.c_float
        L.plh_terrain_update.argtypes = [vp, C.c_double, C.c_double, C.c_double, C.c_float, C.c_float]
        L.plh_terrain_quads.argtypes = [vp, vp, i]
        L.plh_sampler_create.argtypes = [C.c_char_p, vp, i, i]
        L.plh_sampler_z_create.argtypes = [C.c_char_p, vp, i, i]
        L.plh_sampler_z_counts.argtypes = [vp, vp]
        L.plh_ground_height.argtypes = [vp, i]
        L.plh_ground_height.restype = None
        L.plh_terrain_quads_z.argtypes = [vp, vp, i]
        L.plh_sampler_destroy.argtypes = [vp]
        L.plh_sampler_tile_count.argtypes = [vp]
        L.plh_frame_update.argtypes = [vp, vp, vp, i]
        L.plh_test_calls.argtypes = [vp, vp, i]
        _lib = L
    return _lib


def _err():
    return (lib().plh_last_error() or b"").decode()


def _check(rc):
    if rc != 0:
        raise HostError(_err())


class Tile:
    """TileCache::Tile*"""

    def __init__(self, producer, h, level, tx, ty):
        self.producer, self.h, self.level, self.tx, self.ty = producer, h, level, tx, ty

    @property
    def done(self):
        return bool(lib().plh_tile_done(self.h))

    @property
    def slot(self):
        return lib().plh_tile_slot(self.h)

    def download(self):
        shape, dt = self.producer.tile_shape()
        out = np.empty(shape, dt)
        _check(lib().plh_tile_download(self.h, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def minmax(self):
        mm = np.zeros(2, np.float32)
        _check(lib().plh_tile_minmax(self.producer.h, self.h, mm.ctypes.data_as(C.c_void_p)))
        return float(mm[0]), float(mm[1])


class Producer:
    """TileProducer*"""

    def __init__(self, h):
        self.h = h

    def info(self):
        out = (C.c_int * 6)()
        lib().plh_producer_info(self.h, out)
        return dict(id=out[0], border=out[1], gpu=bool(out[2]), tile_size=out[3], referenced=out[4])

    @property
    def type(self):
        return lib().plh_producer_type(self.h).decode()

    @property
    def task_type(self):
        return lib().plh_producer_task_type(self.h).decode()

    def tile_shape(self):
        w = self.info()["tile_size"]
        return {"ElevationProducer": ((w, w, 3), np.float32), "ResidualProducer": ((w, w), np.float32),
                "OrthoProducer": ((w, w, 4), np.uint8), "OrthoCPUProducer": ((w, w, 4), np.uint8), "NormalProducer": None}.get(self.type) or ((w, w, self._norm_channels), np.uint8)

    _norm_channels = 2

    def set_root_quad_size(self, size):
        lib().plh_set_root_quad_size(self.h, size)

    def has_tile(self, level, tx, ty):
        return bool(lib().plh_has_tile(self.h, level, tx, ty))

    def has_children(self, level, tx, ty):
        return bool(lib().plh_has_children(self.h, level, tx, ty))

    def get_tile(self, level, tx, ty, deadline=0):
        h = lib().plh_get_tile(self.h, level, tx, ty, deadline)
        if not h:
            raise HostError(_err())
        return Tile(self, h, level, tx, ty)

    def find_tile(self, level, tx, ty, include_cache=False, done=False):
        h = lib().plh_find_tile(self.h, level, tx, ty, int(include_cache), int(done))
        return Tile(self, h, level, tx, ty) if h else None

    def put_tile(self, tile):
        _check(lib().plh_put_tile(self.h, tile.h))

    def prefetch_tile(self, level, tx, ty):
        rc = lib().plh_prefetch_tile(self.h, level, tx, ty)
        if rc < 0:
            raise HostError(_err())
        return bool(rc)

    def invalidate_tiles(self):
        lib().plh_invalidate_tiles(self.h)

    def invalidate_tile(self, level, tx, ty):
        lib().plh_invalidate_tile(self.h, level, tx, ty)

    def counts(self):
        out = (C.c_ulonglong * 2)()
        _check(lib().plh_producer_counts(self.h, out))
        return int(out[0]), int(out[1])

    @property
    def cache(self):
        return Cache(lib().plh_producer_cache(self.h))

    # ResidualProducer
    def residual_info(self):
        out = (C.c_int * 3)()
        _check(lib().plh_residual_info(self.h, out))
        return dict(min_level=out[0], max_level=out[1], delta=out[2])

    def residual_tile_id(self, level, tx, ty):
        return lib().plh_residual_tile_id(self.h, level, tx, ty)

    def residual_tile_size(self, level):
        return lib().plh_residual_tile_size(self.h, level)


class Cache:
    def __init__(self, h):
        self.h = h

    def stats(self):
        out = (C.c_int * 6)()
        lib().plh_cache_stats(self.h, out)
        return dict(used=out[0], unused=out[1], capacity=out[2], free=out[3], queries=out[4], misses=out[5])

    @property
    def scheduler(self):
        return Scheduler(lib().plh_cache_scheduler(self.h))


class Scheduler:
    def __init__(self, h):
        self.h = h

    def run(self, tiles):
        arr = (C.c_void_p * len(tiles))(*[t.h for t in tiles])
        _check(lib().plh_run(self.h, arr, len(tiles)))

    def stats(self):
        out = (C.c_ulonglong * 4)()
        lib().plh_scheduler_stats(self.h, out)
        return dict(frame=int(out[0]), waves=int(out[1]), tasks=int(out[2]), queued=int(out[3]))


class Scene:
    """The resources of one XML archive (ResourceManager)."""

    def __init__(self, xml, data_dir=".", device=-1):
        self.h = lib().plh_open(xml.encode(), data_dir.encode(), device)
        if not self.h:
            raise HostError(_err())

    def close(self):
        if self.h:
            lib().plh_close(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, exc_type, *a):
        # closing a scene whose tiles are still in use is an assertion failure (TileCache.cpp:113-117);
        # when the block is left by an exception the scene is leaked so that the exception is what is seen
        if exc_type is None:
            self.close()

    def producer(self, name, channels=2):
        h = lib().plh_producer(self.h, name.encode())
        if not h:
            raise HostError(_err())
        p = Producer(h)
        p._norm_channels = channels
        return p

    def cache(self, name):
        h = lib().plh_cache(self.h, name.encode())
        if not h:
            raise HostError(_err())
        return Cache(h)

    def scheduler(self, name):
        h = lib().plh_scheduler(self.h, name.encode())
        if not h:
            raise HostError(_err())
        return Scheduler(h)


class TestScene:
    """CPU-only: a TileCache over a plain TileStorage with a recording producer."""
    __test__ = False

    def __init__(self, capacity, tile_size=8, max_level=30, prefetch_rate=0, prefetch_queue=0):
        self.h = lib().plh_test_scene(capacity, tile_size, max_level, prefetch_rate, prefetch_queue)
        if not self.h:
            raise HostError(_err())
        self.producer = Producer(lib().plh_test_producer(self.h))
        self.cache = Cache(lib().plh_test_cache(self.h))
        self.scheduler = Scheduler(lib().plh_test_scheduler(self.h))

    def calls(self):
        n = lib().plh_test_calls(self.h, None, 0)
        out = (C.c_int * (4 * max(n, 1)))()
        lib().plh_test_calls(self.h, out, n)
        return [tuple(out[4 * k:4 * k + 4]) for k in range(n)]

    def begin_end(self):
        out = (C.c_int * 2)()
        lib().plh_test_begin_end(self.h, out)
        return out[0], out[1]

    def close(self):
        if self.h:
            lib().plh_test_scene_close(self.h)
            self.h = None


class Terrain:
    """TerrainNode: the view-dependent quadtree over the root quad [-size, size]^2."""

    def __init__(self, size, zmin=0.0, zmax=5000.0, split_factor=2.0, max_level=16):
        self.h = lib().plh_terrain_create(size, zmin, zmax, split_factor, max_level)
        self.split_factor = split_factor

    def update(self, x, y, z, split_dist=None, dist_factor=1.0, viewport_width=1024.0, fov=1.3962634):
        if split_dist is None:
            split_dist = lib().plh_split_distance(self.split_factor, viewport_width, fov)
        n = lib().plh_terrain_update(self.h, x, y, z, split_dist, dist_factor)
        if n < 0:
            raise HostError(_err())
        return n

    def quads(self):
        n = lib().plh_terrain_quads(self.h, None, 0)
        out = (C.c_int * (4 * n))()
        lib().plh_terrain_quads(self.h, out, n)
        return [tuple(out[4 * k:4 * k + 4]) for k in range(n)]

    def quads_z(self):
        """pre-order (level, tx, ty, zmin, zmax): the z ranges a TileSamplerZ has read back so far"""
        n = lib().plh_terrain_quads_z(self.h, None, 0)
        out = (C.c_float * (5 * n))()
        lib().plh_terrain_quads_z(self.h, out, n)
        return [(int(out[5 * k]), int(out[5 * k + 1]), int(out[5 * k + 2]), out[5 * k + 3], out[5 * k + 4]) for k in range(n)]

    def close(self):
        if self.h:
            lib().plh_terrain_destroy(self.h)
            self.h = None


class Sampler:
    """TileSampler: holds the tiles of one producer that the terrain's quads need."""

    def __init__(self, name, producer, asynchronous=False, store_parent=True):
        self.h = lib().plh_sampler_create(name.encode(), producer.h, int(asynchronous), int(store_parent))
        if not self.h:
            raise HostError(_err())

    @property
    def tile_count(self):
        return lib().plh_sampler_tile_count(self.h)

    def close(self):
        if self.h:
            lib().plh_sampler_destroy(self.h)
            self.h = None


class SamplerZ(Sampler):
    """TileSamplerZ: a Sampler of elevation tiles that feeds TerrainQuad zmin / zmax and the ground height under the camera"""

    def __init__(self, name, producer, asynchronous=False, store_parent=True):
        self.h = lib().plh_sampler_z_create(name.encode(), producer.h, int(asynchronous), int(store_parent))
        if not self.h:
            raise HostError(_err())

    def counts(self):
        out = (C.c_ulonglong * 3)()
        lib().plh_sampler_z_counts(self.h, out)
        return tuple(out)      # read-backs issued, applied, tiles waiting


def ground_height(reset=False):
    """(TerrainNode::groundHeightAtCamera, nextGroundHeightAtCamera)"""
    out = (C.c_float * 2)()
    lib().plh_ground_height(out, int(reset))
    return out[0], out[1]


def frame_update(scheduler, terrain, samplers):
    """One frame of every sampler against the terrain + Scheduler::run; returns the task count."""
    arr = (C.c_void_p * len(samplers))(*[s.h for s in samplers])
    n = lib().plh_frame_update(scheduler.h, terrain.h, arr, len(samplers))
    if n < 0:
        raise HostError(_err())
    return n


def upsample_variant(prog):
    out = (C.c_int * 2)()
    rc = lib().plh_upsample_variant(prog.encode(), out)
    return None if rc else (bool(out[0]), bool(out[1]))
