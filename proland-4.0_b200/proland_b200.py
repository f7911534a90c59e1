"""ctypes binding of libproland_b200.so (include/proland_b200.h).

This is the Python face of the C ABI used by the tests, bench.py and
__graft_entry__.py.  It holds no arithmetic: everything is computed by the CUDA
library.  There is no fallback: a missing library or a missing GPU raises.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PL_LIB", os.path.join(HERE, "libproland_b200.so"))

PL_OK, PL_ERR_ARG, PL_ERR_POOL_FULL, PL_ERR_CUDA, PL_ERR_CORRUPT, PL_ERR_NO_DEVICE, PL_ERR_IO = range(7)
POOL_ELEV, POOL_NORM2, POOL_NORM4, POOL_RESID_F32, POOL_RESID_I16, POOL_ORTHO = range(6)
NOISE_PLAIN, NOISE_SLOPE = 0, 1
ARITH_EXACT, ARITH_FAST = 0, 1     # pl_norm_scene.arith: the normal pass's arithmetic contract
FILTER_NEAREST, FILTER_LINEAR = 0, 1

# every symbol include/proland_b200.h declares (tests check the .so exports them all)
EXPORTS = [
    "pl_last_error", "pl_abi_version", "pl_ctx_create", "pl_ctx_destroy", "pl_ctx_set_stream",
    "pl_ctx_stream", "pl_sync", "pl_ctx_launch_count", "pl_device_sm_count", "pl_timing_enable",
    "pl_timing_collect", "pl_pool_create",
    "pl_pool_destroy", "pl_pool_capacity", "pl_pool_tile_w", "pl_pool_tile_bytes",
    "pl_pool_slot_bytes", "pl_pool_device_ptr", "pl_pool_download", "pl_pool_upload", "pl_pool_download_range",
    "pl_pool_export", "pl_pool_attach_peers", "pl_pool_push_to_peers",
    "pl_pool_create_shared", "pl_pool_mc_create", "pl_pool_mc_import", "pl_pool_mc_add_device", "pl_pool_mc_bind",
    "pl_noise_init", "pl_noise_select", "pl_cnoise2", "pl_elev_make_req", "pl_elevation_batch",
    "pl_elevation_batch_dev", "pl_elev_stats_download", "pl_elev_stats_range", "pl_elev_stats_readback_begin",
    "pl_elev_stats_readback_end", "pl_elev_zreadback_begin", "pl_elev_stats_readback_ready", "pl_norm_make_req", "pl_normal_batch",
    "pl_normal_batch_dev", "pl_pair_batch", "pl_pair_batch_dev", "pl_pair_batch_ids", "pl_produce_levels", "pl_make_tile_ids_range", "pl_produce_range", "pl_make_requests_range",
    "pl_debug_download_requests", "pl_debug_force_generic", "pl_debug_no_fuse", "pl_debug_no_slim", "pl_debug_inflate_path", "pl_debug_stage_ring", "pl_debug_fpexact",
    "pl_residual_decode_batch", "pl_blobs_create", "pl_blobs_destroy", "pl_residual_decode_stored", "pl_residual_upsample", "pl_residual_encode_batch", "pl_residual_write_file",
    "pl_height_cube_create", "pl_height_cube_from_latlon", "pl_height_cube_from_plane", "pl_debug_height_unsure", "pl_height_cube_download", "pl_height_cube_destroy", "pl_height_tiles",
    "pl_ortho_noise_init", "pl_ortho_noise_host", "pl_ortho_make_req", "pl_ortho_make_requests_range", "pl_ortho_batch", "pl_ortho_batch_dev", "pl_ortho_decode_batch", "pl_ortho_produce_range",
]


class PlError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("pl error %d: %s" % (code, msg))
        self.code = code


class ElevScene(C.Structure):
    _fields_ = [("tile_w", C.c_int32), ("grid", C.c_int32), ("flip", C.c_int32),
                ("noise_mode", C.c_int32), ("no_clamp", C.c_int32), ("want_stats", C.c_int32),
                ("resid_scale", C.c_float), ("pad_", C.c_int32)]


class ElevReq(C.Structure):
    _fields_ = [("out_slot", C.c_int32), ("parent_slot", C.c_int32), ("resid_slot", C.c_int32),
                ("dx", C.c_int32), ("dy", C.c_int32), ("rx", C.c_int32), ("ry", C.c_int32),
                ("noise_r", C.c_int32), ("noise_l", C.c_int32), ("rs", C.c_float),
                ("pixel_size", C.c_float), ("level", C.c_int32), ("tx", C.c_int32),
                ("ty", C.c_int32), ("pad_", C.c_int32 * 2)]


class NormScene(C.Structure):
    _fields_ = [("tile_w", C.c_int32), ("grid", C.c_int32), ("elev_border", C.c_int32),
                ("elev_filter", C.c_int32), ("parent_filter", C.c_int32), ("sphere", C.c_int32),
                ("arith", C.c_int32), ("pad_", C.c_int32)]


class NormReq(C.Structure):
    _fields_ = [("out_slot", C.c_int32), ("elev_slot", C.c_int32), ("parent_slot", C.c_int32),
                ("ptx", C.c_int32), ("pty", C.c_int32), ("level", C.c_int32),
                ("deform", C.c_float * 4), ("corners", C.c_float * 12),
                ("verticals", C.c_float * 12), ("norms", C.c_float * 4), ("w2t", C.c_float * 9),
                ("p2t", C.c_float * 9), ("smooth", C.c_float), ("pad_", C.c_int32 * 3)]


class SweepScene(C.Structure):
    _fields_ = [("elev", ElevScene), ("norm", NormScene), ("root_quad_size", C.c_float),
                ("face", C.c_int32), ("n_amp", C.c_int32), ("pad_", C.c_int32),
                ("noise_amp", C.c_float * 32)]


class OrthoScene(C.Structure):
    _fields_ = [("tile_w", C.c_int32), ("channels", C.c_int32), ("hsv", C.c_int32), ("face", C.c_int32),
                ("scale", C.c_float), ("noise_color", C.c_float * 4), ("root_noise_color", C.c_float * 4),
                ("n_amp", C.c_int32), ("max_level", C.c_int32), ("out_channels", C.c_int32), ("noise_amp", C.c_float * 32)]


ORTHO_REQ_DTYPE = np.dtype([("out_slot", "i4"), ("parent_slot", "i4"), ("resid_slot", "i4"), ("dx", "i4"), ("dy", "i4"),
                            ("noise_r", "i4"), ("noise_l", "i4"), ("level", "i4"), ("noise_color", "f4", (4,)),
                            ("tx", "i4"), ("ty", "i4"), ("pad_", "i4", (2,))])
assert ORTHO_REQ_DTYPE.itemsize == 64 and C.sizeof(OrthoScene) == 192

assert C.sizeof(ElevReq) == 64 and C.sizeof(NormReq) == 240 and C.sizeof(ElevScene) == 32

ELEV_REQ_DTYPE = np.dtype([("out_slot", "i4"), ("parent_slot", "i4"), ("resid_slot", "i4"),
                           ("dx", "i4"), ("dy", "i4"), ("rx", "i4"), ("ry", "i4"),
                           ("noise_r", "i4"), ("noise_l", "i4"), ("rs", "f4"),
                           ("pixel_size", "f4"), ("level", "i4"), ("tx", "i4"), ("ty", "i4"),
                           ("pad_", "i4", (2,))])
NORM_REQ_DTYPE = np.dtype([("out_slot", "i4"), ("elev_slot", "i4"), ("parent_slot", "i4"),
                           ("ptx", "i4"), ("pty", "i4"), ("level", "i4"), ("deform", "f4", (4,)),
                           ("corners", "f4", (12,)), ("verticals", "f4", (12,)),
                           ("norms", "f4", (4,)), ("w2t", "f4", (9,)), ("p2t", "f4", (9,)),
                           ("smooth", "f4"), ("pad_", "i4", (3,))])
RESID_ENC_DTYPE = np.dtype([("tile_slot", "i4"), ("parent_slot", "i4"), ("approx_slot", "i4"), ("resid_slot", "i4"),
                            ("tile_size", "i4"), ("tx", "i4"), ("ty", "i4"), ("pad_", "i4")])
TILE_ID_DTYPE = np.dtype([("level", "i4"), ("tx", "i4"), ("ty", "i4"), ("elev_slot", "i4"), ("parent_slot", "i4"),
                          ("resid_slot", "i4"), ("norm_slot", "i4"), ("pad_", "i4")])
assert ELEV_REQ_DTYPE.itemsize == 64 and NORM_REQ_DTYPE.itemsize == 240 and RESID_ENC_DTYPE.itemsize == 32
LEVEL_RANGE_DTYPE = np.dtype([("level", "i4"), ("n", "i4"), ("morton0", "u8"), ("out_slot0", "i4"), ("parent_slot0", "i4"),
                              ("parent_morton0", "u8")])
assert TILE_ID_DTYPE.itemsize == 32 and LEVEL_RANGE_DTYPE.itemsize == 32
HEIGHT_REQ_DTYPE = np.dtype([("face", "i4"), ("level", "i4"), ("tx", "i4"), ("ty", "i4"), ("out_slot", "i4"), ("pad_", "i4", (3,))])
assert HEIGHT_REQ_DTYPE.itemsize == 32

_lib = None


def build():
    """Compile libproland_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.check_call(["make", "-s", "-C", HERE])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PlError(PL_ERR_NO_DEVICE, "%s is missing: run `make -C %s` (there is no fallback path)"
                          % (LIB_PATH, HERE))
        L = C.CDLL(LIB_PATH)
        L.pl_last_error.restype = C.c_char_p
        L.pl_ctx_stream.restype = C.c_void_p
        L.pl_ctx_launch_count.restype = C.c_uint64
        L.pl_pool_tile_bytes.restype = C.c_size_t
        L.pl_pool_slot_bytes.restype = C.c_size_t
        L.pl_pool_device_ptr.restype = C.c_void_p
        L.pl_cnoise2.restype = C.c_float
        L.pl_cnoise2.argtypes = [C.c_float, C.c_float]
        L.pl_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.pl_ctx_destroy.argtypes = [C.c_void_p]
        L.pl_ctx_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        L.pl_ctx_stream.argtypes = [C.c_void_p]
        L.pl_sync.argtypes = [C.c_void_p]
        L.pl_ctx_launch_count.argtypes = [C.c_void_p]
        L.pl_device_sm_count.argtypes = [C.c_void_p]
        L.pl_pool_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.pl_pool_destroy.argtypes = [C.c_void_p]
        for f in (L.pl_pool_capacity, L.pl_pool_tile_w, L.pl_pool_tile_bytes, L.pl_pool_slot_bytes,
                  L.pl_pool_device_ptr):
            f.argtypes = [C.c_void_p]
        L.pl_pool_export.argtypes = [C.c_void_p, C.c_void_p]
        L.pl_pool_attach_peers.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.pl_pool_push_to_peers.argtypes = [C.c_void_p, C.c_int]
        L.pl_pool_create_shared.argtypes = L.pl_pool_create.argtypes
        L.pl_pool_mc_create.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.pl_pool_mc_import.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.pl_pool_mc_add_device.argtypes = [C.c_void_p]
        L.pl_pool_mc_bind.argtypes = [C.c_void_p]
        L.pl_pool_download.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        L.pl_pool_download_range.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
        L.pl_pool_upload.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        L.pl_noise_init.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.pl_noise_select.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int),
                                      C.POINTER(C.c_int)]
        L.pl_noise_select.restype = None
        L.pl_elev_make_req.argtypes = [C.c_int, C.c_float, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                       C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.pl_elev_make_req.restype = None
        L.pl_elevation_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                         C.c_void_p]
        L.pl_elevation_batch_dev.argtypes = L.pl_elevation_batch.argtypes
        L.pl_elev_stats_download.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.pl_elev_stats_range.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.pl_elev_stats_readback_begin.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.pl_elev_stats_readback_end.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.pl_norm_make_req.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p]
        L.pl_norm_make_req.restype = None
        L.pl_normal_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                      C.c_void_p]
        L.pl_normal_batch_dev.argtypes = L.pl_normal_batch.argtypes
        L.pl_pair_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_int, C.c_void_p, C.c_void_p]
        L.pl_pair_batch_dev.argtypes = L.pl_pair_batch.argtypes
        L.pl_pair_batch_ids.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.pl_produce_levels.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.pl_make_tile_ids_range.argtypes = [C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_void_p]
        L.pl_produce_range.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                       C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_uint64]
        L.pl_make_requests_range.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int,
                                             C.c_uint64, C.c_void_p, C.c_void_p, C.c_int]
        L.pl_debug_download_requests.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.pl_debug_force_generic.argtypes = [C.c_void_p, C.c_int]
        L.pl_debug_no_fuse.argtypes = [C.c_void_p, C.c_int]
        L.pl_debug_no_slim.argtypes = [C.c_void_p, C.c_int]
        L.pl_debug_inflate_path.argtypes = [C.c_void_p, C.c_int]
        L.pl_debug_stage_ring.argtypes = [C.c_void_p, C.c_size_t]
        L.pl_debug_fpexact.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.pl_residual_decode_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_float]
        L.pl_blobs_create.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]
        L.pl_blobs_destroy.argtypes = [C.c_void_p]
        L.pl_blobs_destroy.restype = None
        L.pl_residual_decode_stored.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_float]
        L.pl_residual_upsample.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.pl_residual_write_file.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                             C.c_void_p, C.c_void_p, C.c_int]
        L.pl_residual_encode_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                               C.c_void_p, C.c_void_p]
        L.pl_height_cube_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
        L.pl_height_cube_from_latlon.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.pl_height_cube_from_plane.argtypes = L.pl_height_cube_from_latlon.argtypes
        L.pl_debug_height_unsure.argtypes = [C.c_void_p]
        L.pl_debug_height_unsure.restype = C.c_uint64
        L.pl_height_cube_download.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.pl_height_cube_destroy.argtypes = [C.c_void_p]
        L.pl_height_cube_destroy.restype = None
        L.pl_height_tiles.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p]
        L.pl_ortho_noise_init.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.pl_ortho_noise_host.argtypes = [C.c_int, C.c_void_p]
        L.pl_ortho_make_req.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.pl_ortho_make_req.restype = None
        L.pl_ortho_make_requests_range.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_uint64,
                                                   C.c_void_p, C.c_int]
        L.pl_ortho_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.pl_ortho_batch_dev.argtypes = L.pl_ortho_batch.argtypes
        L.pl_ortho_produce_range.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int,
                                             C.c_uint64]
        L.pl_ortho_decode_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.POINTER(C.c_int)]
        L.pl_timing_enable.argtypes = [C.c_void_p, C.c_int]
        L.pl_timing_collect.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def check(rc):
    if rc != PL_OK:
        raise PlError(rc, lib().pl_last_error().decode("utf-8", "replace"))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


# ----------------------------------------------------------------- host helpers

def cnoise2(x, y):
    return lib().pl_cnoise2(x, y)


def noise_select(level, tx, ty, face):
    r, l = C.c_int(), C.c_int()
    lib().pl_noise_select(level, tx, ty, face, C.byref(r), C.byref(l))
    return r.value, l.value


def elev_make_reqs(tiles, *, tile_w=101, root_quad_size=100000.0, noise_amp=(), face=0,
                   resid_tile_w=0, has_resid=None):
    """tiles: iterable of (level, tx, ty) -> structured array of pl_elev_req (slots = -1)."""
    tiles = list(tiles)
    out = np.zeros(len(tiles), ELEV_REQ_DTYPE)
    amp = np.asarray(noise_amp, np.float32)
    L = lib()
    for i, (level, tx, ty) in enumerate(tiles):
        hr = 0 if has_resid is None else int(bool(has_resid[i]))
        L.pl_elev_make_req(tile_w, C.c_float(root_quad_size), _ptr(amp), len(amp), face, level, tx,
                           ty, resid_tile_w, hr, C.c_void_p(out.ctypes.data + 64 * i))
    return out


def norm_make_reqs(tiles, scene, *, root_quad_size=100000.0, components=2):
    tiles = list(tiles)
    out = np.zeros(len(tiles), NORM_REQ_DTYPE)
    L = lib()
    for i, (level, tx, ty) in enumerate(tiles):
        L.pl_norm_make_req(C.byref(scene), C.c_double(root_quad_size), components, level, tx, ty,
                           C.c_void_p(out.ctypes.data + 240 * i))
    return out


def sweep_scene(*, noise_amp, face=0, root_quad_size=100000.0, tile_w=101, grid_size=24, flip=0,
                noise_mode=NOISE_SLOPE, no_clamp=0, want_stats=0, sphere=0,
                elev_filter=FILTER_LINEAR, arith=ARITH_EXACT):
    s = SweepScene()
    s.elev = elev_scene(tile_w, grid_size, flip, noise_mode, no_clamp, want_stats)
    s.norm = norm_scene(tile_w - 4, grid_size, 2, elev_filter, FILTER_LINEAR, sphere, arith)
    s.root_quad_size = root_quad_size
    s.face = face
    s.n_amp = len(noise_amp)
    for i, a in enumerate(noise_amp):
        s.noise_amp[i] = a
    return s


def make_requests_range(scene, level, morton0, n, out_slot0=0, parent_slot0=0, parent_morton0=0,
                        nthreads=0, normals=True, out=None):
    """Host-built requests of a Morton range (all hardware threads by default).  out = (e, q):
    preallocated request arrays of at least n entries to fill instead of allocating new ones."""
    if out is not None:
        e, q = out[0][:n], (out[1][:n] if normals else None)
    else:
        e = np.zeros(n, ELEV_REQ_DTYPE)
        q = np.zeros(n, NORM_REQ_DTYPE) if normals else None
    check(lib().pl_make_requests_range(C.byref(scene), level, morton0, n, out_slot0, parent_slot0,
                                       parent_morton0, _ptr(e), _ptr(q) if normals else None, nthreads))
    return e, q


def make_tile_ids_range(level, morton0, n, out_slot0=0, parent_slot0=0, parent_morton0=0, out=None):
    """tile identities of a Morton range (pl_make_tile_ids_range); out: a preallocated TILE_ID_DTYPE array"""
    ids = out[:n] if out is not None else np.zeros(n, TILE_ID_DTYPE)
    check(lib().pl_make_tile_ids_range(level, morton0, n, out_slot0, parent_slot0, parent_morton0, _ptr(ids)))
    return ids


def residual_write_file(path, tiles, min_level, max_level, tile_size, root=(0, 0, 0), scale=1.0, zlib_level=-1):
    """tiles: {tile id: (w, w) int16} -> a residual file in the reference's format (pl_residual_write_file)"""
    ids = sorted(tiles)
    flat = [np.ascontiguousarray(tiles[i], np.int16).ravel() for i in ids]
    assert ids == list(range(len(ids)))
    offs = np.concatenate([[0], np.cumsum([len(a) for a in flat[:-1]])]).astype(np.uint64)
    buf = np.concatenate(flat)
    check(lib().pl_residual_write_file(os.fsencode(path), min_level, max_level, tile_size, root[0], root[1], root[2],
                                       C.c_float(scale), _ptr(buf), _ptr(offs), zlib_level))


def morton_encode(tx, ty):
    m = 0
    for b in range(24):
        m |= ((tx >> b) & 1) << (2 * b) | ((ty >> b) & 1) << (2 * b + 1)
    return m


def morton_decode(m):
    tx = ty = 0
    for b in range(24):
        tx |= ((m >> (2 * b)) & 1) << b
        ty |= ((m >> (2 * b + 1)) & 1) << b
    return tx, ty


# ------------------------------------------------------------------- context

class Pool:
    def __init__(self, ctx, kind, tile_w, capacity, shared=False):
        self.ctx = ctx
        self.kind = kind
        h = C.c_void_p()
        # shared: the memory comes from the VMM allocator and can be bound to an NVLink multicast object (pl_pool_mc_*)
        check((lib().pl_pool_create_shared if shared else lib().pl_pool_create)(ctx.h, kind, tile_w, capacity, C.byref(h)))
        self.h = h
        self.tile_w = tile_w
        self.capacity = capacity
        self.tile_bytes = lib().pl_pool_tile_bytes(h)
        self.slot_bytes = lib().pl_pool_slot_bytes(h)

    def close(self):
        if self.h:
            lib().pl_pool_destroy(self.h)
            self.h = None

    def export(self):
        """the 64-byte CUDA IPC handle of the pool's memory (pl_pool_export)"""
        h = np.zeros(64, np.uint8)
        check(lib().pl_pool_export(self.h, _ptr(h)))
        return h

    def attach_peers(self, handles, self_rank):
        """handles: (world, 64) uint8, the exported handles of all ranks in rank order"""
        hs = np.ascontiguousarray(handles, np.uint8)
        check(lib().pl_pool_attach_peers(self.h, len(hs), _ptr(hs), self_rank))

    def push_to_peers(self, on=True):
        """True / 1: one unicast store per attached peer; 2: one store through the multicast mapping; False: off"""
        check(lib().pl_pool_push_to_peers(self.h, int(on)))

    def mc_create(self, n_devices):
        """rank 0: the multicast object of the group -> a file descriptor to hand to the other ranks (pl_pool_mc_create)"""
        fd = C.c_int(-1)
        check(lib().pl_pool_mc_create(self.h, n_devices, C.byref(fd)))
        return fd.value

    def mc_import(self, fd, n_devices):
        check(lib().pl_pool_mc_import(self.h, fd, n_devices))

    def mc_add_device(self):
        check(lib().pl_pool_mc_add_device(self.h))

    def mc_bind(self):
        check(lib().pl_pool_mc_bind(self.h))

    def _shape_dtype(self):
        W = self.tile_w
        return {POOL_ELEV: ((W, W, 3), np.float32), POOL_NORM2: ((W, W, 2), np.uint8),
                POOL_NORM4: ((W, W, 4), np.uint8), POOL_RESID_F32: ((W, W), np.float32),
                POOL_RESID_I16: ((W, W), np.int16), POOL_ORTHO: ((W, W, 4), np.uint8)}[self.kind]

    def download(self, slot):
        shape, dt = self._shape_dtype()
        out = np.empty(shape, dt)
        check(lib().pl_pool_download(self.h, slot, _ptr(out), out.nbytes))
        return out

    def download_range(self, slot0, n):
        """n consecutive slots in one copy (pl_pool_download_range) -> (n,) + the tile shape"""
        shape, dt = self._shape_dtype()
        out = np.empty((n,) + shape, dt)
        check(lib().pl_pool_download_range(self.h, slot0, n, _ptr(out), out.nbytes))
        return out

    def upload(self, slot, arr):
        shape, dt = self._shape_dtype()
        a = np.ascontiguousarray(arr, dt).reshape(shape)
        check(lib().pl_pool_upload(self.h, slot, _ptr(a), a.nbytes))


class Context:
    def __init__(self, device=0):
        h = C.c_void_p()
        check(lib().pl_ctx_create(device, C.byref(h)))
        self.h = h
        self.device = device
        self._pools = []

    def close(self):
        for p in self._pools:
            p.close()
        self._pools = []
        if self.h:
            lib().pl_ctx_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def pool(self, kind, tile_w, capacity, shared=False):
        p = Pool(self, kind, tile_w, capacity, shared)
        self._pools.append(p)
        return p

    def set_stream(self, cuda_stream):
        check(lib().pl_ctx_set_stream(self.h, C.c_void_p(cuda_stream)))

    @property
    def stream(self):
        return lib().pl_ctx_stream(self.h)

    def sync(self):
        check(lib().pl_sync(self.h))

    @property
    def launches(self):
        return lib().pl_ctx_launch_count(self.h)

    @property
    def sm_count(self):
        return lib().pl_device_sm_count(self.h)

    def noise_init(self, tile_w=101):
        out = np.empty((6, tile_w, tile_w), np.float32)
        check(lib().pl_noise_init(self.h, tile_w, _ptr(out)))
        return out

    def elevation_batch(self, scene, elev, reqs, resid=None):
        reqs = np.ascontiguousarray(reqs, ELEV_REQ_DTYPE)
        check(lib().pl_elevation_batch(self.h, C.byref(scene), elev.h, resid.h if resid else None,
                                       len(reqs), _ptr(reqs)))

    def elevation_batch_dev(self, scene, elev, n, dev_ptr, resid=None):
        check(lib().pl_elevation_batch_dev(self.h, C.byref(scene), elev.h,
                                           resid.h if resid else None, n, C.c_void_p(dev_ptr)))

    def elev_stats(self, elev, slots):
        slots = np.ascontiguousarray(slots, np.int32)
        out = np.empty((len(slots), 2), np.float32)
        check(lib().pl_elev_stats_download(self.h, elev.h, len(slots), _ptr(slots), _ptr(out)))
        return out

    def elev_stats_range(self, elev, slot0, n):
        out = np.empty((n, 2), np.float32)
        check(lib().pl_elev_stats_range(self.h, elev.h, slot0, n, _ptr(out)))
        return out

    def elev_stats_readback_begin(self, elev, slot0, n):
        """enqueue the read-back of n (zmin, zmax) pairs; -> ticket for elev_stats_readback_end"""
        t = C.c_int()
        check(lib().pl_elev_stats_readback_begin(self.h, elev.h, slot0, n, C.byref(t)))
        return (t.value, n)

    def elev_stats_readback_end(self, ticket):
        out = np.empty((ticket[1], 2), np.float32)
        check(lib().pl_elev_stats_readback_end(self.h, ticket[0], _ptr(out)))
        return out

    def normal_batch(self, scene, norm, elev, reqs):
        reqs = np.ascontiguousarray(reqs, NORM_REQ_DTYPE)
        check(lib().pl_normal_batch(self.h, C.byref(scene), norm.h, elev.h, len(reqs), _ptr(reqs)))

    def pair_batch(self, escene, nscene, elev, norm, ereqs, nreqs, resid=None):
        """elevation + normal tile pairs from HOST request arrays, one fused kernel (pl_pair_batch)."""
        ereqs = np.ascontiguousarray(ereqs, ELEV_REQ_DTYPE)
        nreqs = np.ascontiguousarray(nreqs, NORM_REQ_DTYPE)
        assert len(ereqs) == len(nreqs)
        check(lib().pl_pair_batch(self.h, C.byref(escene), C.byref(nscene), elev.h, norm.h,
                                  resid.h if resid else None, len(ereqs), _ptr(ereqs), _ptr(nreqs)))

    def produce_levels(self, scene, elev, norm, ranges):
        """consecutive levels of a subtree in one launch (pl_produce_levels); ranges: (level, morton0, n, out_slot0,
        parent_slot0, parent_morton0) tuples, each laid out like a produce_range call"""
        arr = np.zeros(len(ranges), LEVEL_RANGE_DTYPE)
        for k, (level, m0, n, s0, p0, pm0) in enumerate(ranges):
            arr[k] = (level, n, m0, s0, p0, pm0)
        check(lib().pl_produce_levels(self.h, C.byref(scene), elev.h, norm.h, len(arr), _ptr(arr)))

    def pair_batch_ids(self, scene, elev, norm, ids, resid=None):
        """elevation + normal tile pairs from 32-byte tile identities (HOST array); the uniforms are expanded on the
        device (pl_pair_batch_ids)."""
        ids = np.ascontiguousarray(ids, TILE_ID_DTYPE)
        check(lib().pl_pair_batch_ids(self.h, C.byref(scene), elev.h, norm.h, resid.h if resid else None, len(ids), _ptr(ids)))

    def residual_encode(self, heights, approx, resid, reqs):
        """one level of the residual-pyramid builder (pl_residual_encode_batch);
        -> (max |residual|, max |heights - approximation|) per tile"""
        reqs = np.ascontiguousarray(reqs, RESID_ENC_DTYPE)
        mr = np.empty(len(reqs), np.float32)
        me = np.empty(len(reqs), np.float32)
        check(lib().pl_residual_encode_batch(self.h, heights.h, approx.h, resid.h, len(reqs), _ptr(reqs),
                                             _ptr(mr), _ptr(me)))
        return mr, me

    def height_cube(self, faces):
        """the base-level grids of a residual builder on the device: faces = 1 (flat DEM) or 6 (setCube order hm1..hm6)
        int16 arrays of (B + 1, B + 1) samples (pl_height_cube_create)"""
        faces = [np.ascontiguousarray(f, np.int16) for f in faces]
        B = faces[0].shape[0] - 1
        assert all(f.shape == (B + 1, B + 1) for f in faces)
        ptrs = (C.c_void_p * len(faces))(*[f.ctypes.data for f in faces])
        h = C.c_void_p()
        check(lib().pl_height_cube_create(self.h, B, len(faces), ptrs, C.byref(h)))
        return HeightCube(self, h, B, len(faces))

    def height_cube_from_latlon(self, src, base_size):
        """SphericalHeightFunction on the device: six base grids from an equirectangular float map (pl_height_cube_from_latlon)"""
        src = np.ascontiguousarray(src, np.float32)
        h = C.c_void_p()
        check(lib().pl_height_cube_from_latlon(self.h, base_size, _ptr(src), src.shape[1], src.shape[0], C.byref(h)))
        return HeightCube(self, h, base_size, 6)

    def height_cube_from_plane(self, src, base_size):
        """PlaneHeightFunction on the device: the base grid of a flat DEM (pl_height_cube_from_plane)"""
        src = np.ascontiguousarray(src, np.float32)
        h = C.c_void_p()
        check(lib().pl_height_cube_from_plane(self.h, base_size, _ptr(src), src.shape[1], src.shape[0], C.byref(h)))
        return HeightCube(self, h, base_size, 1)

    def normal_batch_dev(self, scene, norm, elev, n, dev_ptr):
        check(lib().pl_normal_batch_dev(self.h, C.byref(scene), norm.h, elev.h, n,
                                        C.c_void_p(dev_ptr)))


class HeightCube:
    """pl_height_cube: the resident base level of HeightMipmap (one flat DEM or the six faces of a cube)"""

    def __init__(self, ctx, h, B, nfaces):
        self.ctx, self.h, self.B, self.nfaces = ctx, h, B, nfaces

    def download(self, face):
        out = np.empty((self.B + 1, self.B + 1), np.int16)
        check(lib().pl_height_cube_download(self.ctx.h, self.h, face, _ptr(out)))
        return out

    def tiles(self, heights, top_level_size, tile_size, reqs, scale=1.0):
        """HeightMipmap::getTile for a batch of (face, level, tx, ty, out_slot) (pl_height_tiles)"""
        reqs = np.ascontiguousarray(reqs, HEIGHT_REQ_DTYPE)
        check(lib().pl_height_tiles(self.ctx.h, self.h, heights.h, top_level_size, tile_size, scale, len(reqs), _ptr(reqs)))

    def close(self):
        if self.h:
            lib().pl_height_cube_destroy(self.h)
            self.h = None


def _produce_range(self, scene, elev, norm, level, morton0, n, out_slot0, parent_slot0=0,
                   parent_morton0=0):
    check(lib().pl_produce_range(self.h, C.byref(scene), elev.h, norm.h if norm else None, level,
                                 morton0, n, out_slot0, parent_slot0, parent_morton0))


def _last_requests(self, n):
    e = np.zeros(n, ELEV_REQ_DTYPE)
    q = np.zeros(n, NORM_REQ_DTYPE)
    check(lib().pl_debug_download_requests(self.h, n, _ptr(e), _ptr(q)))
    return e, q


def _timing_enable(self, on=True):
    check(lib().pl_timing_enable(self.h, int(on)))


def _timing_collect(self):
    """-> {kernel name: (total ms, launches, tiles)} since the last collect (synchronises)."""
    ms = np.zeros(6, np.float64)
    cnt = np.zeros(6, np.uint64)
    tiles = np.zeros(6, np.uint64)
    check(lib().pl_timing_collect(self.h, _ptr(ms), _ptr(cnt), _ptr(tiles)))
    names = ("elevation", "normal", "genreq", "residual", "pair", "ortho")
    return {n: (float(ms[i]), int(cnt[i]), int(tiles[i])) for i, n in enumerate(names)}


def _force_generic(self, on=True):
    check(lib().pl_debug_force_generic(self.h, int(on)))


def _no_fuse(self, on=True):
    """pl_produce_range launches the elevation and normal passes separately (same results)."""
    check(lib().pl_debug_no_fuse(self.h, int(on)))


def _no_slim(self, on=True):
    """the fused kernel keeps its 3-CTA layout even when every tile takes the register form (same results)."""
    check(lib().pl_debug_no_slim(self.h, int(on)))


def _fpexact(self, a, b):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    out = np.empty((6, len(a)), np.float32)
    check(lib().pl_debug_fpexact(self.h, len(a), _ptr(a), _ptr(b), _ptr(out)))
    return out


def _residual_decode(self, pool, blobs, widths, out_slots, add_slots=None, scale=1.0):
    """blobs: list of bytes (one TIFF blob per tile)."""
    n = len(blobs)
    buf = np.frombuffer(b"".join(blobs), np.uint8)
    sizes = np.array([len(b) for b in blobs], np.uint32)
    offs = np.concatenate([[0], np.cumsum(sizes[:-1], dtype=np.uint64)]).astype(np.uint64)
    widths = np.ascontiguousarray(widths, np.int32)
    out_slots = np.ascontiguousarray(out_slots, np.int32)
    add = np.ascontiguousarray(add_slots, np.int32) if add_slots is not None else None
    check(lib().pl_residual_decode_batch(self.h, pool.h, n, _ptr(buf), _ptr(offs), _ptr(sizes), _ptr(widths),
                                         _ptr(out_slots), _ptr(add) if add is not None else None,
                                         C.c_float(scale)))


class Blobs:
    """a residual archive resident in device memory (pl_blobs_create); keeps the host copy for IFD parsing"""

    def __init__(self, ctx, data):
        self.ctx = ctx
        self.host = np.frombuffer(bytes(data), np.uint8)
        h = C.c_void_p()
        check(lib().pl_blobs_create(ctx.h, _ptr(self.host), len(self.host), C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            lib().pl_blobs_destroy(self.h)
            self.h = None


def _residual_decode_stored(self, pool, store, offsets, sizes, widths, out_slots, add_slots=None, scale=1.0):
    """tiles of a device-resident archive (pl_residual_decode_stored): offsets / sizes of the TIFF blobs inside it"""
    offs = np.ascontiguousarray(offsets, np.uint64)
    sizes = np.ascontiguousarray(sizes, np.uint32)
    widths = np.ascontiguousarray(widths, np.int32)
    out_slots = np.ascontiguousarray(out_slots, np.int32)
    add = np.ascontiguousarray(add_slots, np.int32) if add_slots is not None else None
    check(lib().pl_residual_decode_stored(self.h, pool.h, store.h, _ptr(store.host), len(offs), _ptr(offs), _ptr(sizes),
                                          _ptr(widths), _ptr(out_slots), _ptr(add) if add is not None else None,
                                          C.c_float(scale)))


def _residual_upsample(self, pool, src_slot, dst_slot, tile_size, tx=0, ty=0):
    check(lib().pl_residual_upsample(self.h, pool.h, src_slot, dst_slot, tile_size, tx, ty))


def ortho_scene(*, tile_w=196, channels=4, hsv=0, face=1, scale=2.0, cnoise=(255, 255, 255, 255),
                rnoise=(127.5, 127.5, 127.5, 127.5), noise_amp=(), max_level=-1, out_channels=0):
    """The orthoProducer resource (OrthoProducer.cpp:440-512): cnoise / rnoise are the XML's 0..255 values
    (divided by 255 as a float, `(float) atof(..) / 255`); a 3-value list keeps the default of the 4th."""
    s = OrthoScene()
    s.tile_w, s.channels, s.hsv, s.face, s.scale = tile_w, channels, int(hsv), face, scale
    nc = [np.float32(1.0)] * 4
    rc = [np.float32(0.5)] * 4
    for i, v in enumerate(cnoise):
        nc[i] = np.float32(v) / np.float32(255)
    for i, v in enumerate(rnoise):
        rc[i] = np.float32(v) / np.float32(255)
    for i in range(4):
        s.noise_color[i] = nc[i]
        s.root_noise_color[i] = rc[i]
    s.n_amp = len(noise_amp)
    s.max_level = max_level
    s.out_channels = out_channels
    for i, a in enumerate(noise_amp):
        s.noise_amp[i] = a
    return s


def ortho_noise_host(tile_w=196):
    out = np.empty((6, tile_w, tile_w, 4), np.uint8)
    check(lib().pl_ortho_noise_host(tile_w, _ptr(out)))
    return out


def ortho_make_reqs(scene, tiles, has_resid=None):
    """tiles: iterable of (level, tx, ty) -> structured array of pl_ortho_req (slots = -1)."""
    tiles = list(tiles)
    out = np.zeros(len(tiles), ORTHO_REQ_DTYPE)
    L = lib()
    for i, (level, tx, ty) in enumerate(tiles):
        hr = 0 if has_resid is None else int(bool(has_resid[i]))
        L.pl_ortho_make_req(C.byref(scene), level, tx, ty, hr, C.c_void_p(out.ctypes.data + 64 * i))
    return out


def ortho_make_requests_range(scene, level, morton0, n, out_slot0=0, parent_slot0=0, parent_morton0=0, nthreads=0,
                              out=None):
    q = out[:n] if out is not None else np.zeros(n, ORTHO_REQ_DTYPE)
    check(lib().pl_ortho_make_requests_range(C.byref(scene), level, morton0, n, out_slot0, parent_slot0,
                                             parent_morton0, _ptr(q), nthreads))
    return q


def _ortho_noise_init(self, tile_w=196, want_host=False):
    out = np.empty((6, tile_w, tile_w, 4), np.uint8) if want_host else None
    check(lib().pl_ortho_noise_init(self.h, tile_w, _ptr(out) if want_host else None))
    return out


def _ortho_batch(self, scene, ortho, resid, reqs):
    reqs = np.ascontiguousarray(reqs, ORTHO_REQ_DTYPE)
    check(lib().pl_ortho_batch(self.h, C.byref(scene), ortho.h, resid.h if resid is not None else None, len(reqs),
                               _ptr(reqs)))


def _ortho_decode(self, pool, blobs, out_slots):
    """blobs: list of bytes (one TIFF blob per byte tile) -> the blobs' channel count"""
    n = len(blobs)
    buf = np.frombuffer(b"".join(blobs), np.uint8)
    sizes = np.array([len(b) for b in blobs], np.uint32)
    offs = np.concatenate([[0], np.cumsum(sizes[:-1], dtype=np.uint64)]).astype(np.uint64)
    out_slots = np.ascontiguousarray(out_slots, np.int32)
    ch = C.c_int(0)
    check(lib().pl_ortho_decode_batch(self.h, pool.h, n, _ptr(buf), _ptr(offs), _ptr(sizes), _ptr(out_slots), C.byref(ch)))
    return ch.value


def _ortho_produce_range(self, scene, ortho, level, morton0, n, out_slot0, parent_slot0=0, parent_morton0=0):
    check(lib().pl_ortho_produce_range(self.h, C.byref(scene), ortho.h, level, morton0, n, out_slot0, parent_slot0,
                                       parent_morton0))


SLOT_SCRATCH = -2
Context.ortho_produce_range = _ortho_produce_range
Context.ortho_decode = _ortho_decode
Context.ortho_noise_init = _ortho_noise_init
Context.ortho_batch = _ortho_batch
Context.residual_decode = _residual_decode
Context.residual_upsample = _residual_upsample
Context.force_generic = _force_generic
Context.residual_decode_stored = _residual_decode_stored
Context.blobs = lambda self, data: Blobs(self, data)
Context.no_fuse = _no_fuse
Context.no_slim = _no_slim


def _inflate_path(self, path=0):
    """0: decoder chosen by the batch size, 1: warp-per-stream kernel, 2: tokenizer + resolver kernels"""
    check(lib().pl_debug_inflate_path(self.h, int(path)))


Context.inflate_path = _inflate_path
Context.stage_ring = lambda self, min_bytes: check(lib().pl_debug_stage_ring(self.h, min_bytes))
Context.fpexact = _fpexact
Context.timing_enable = _timing_enable
Context.timing_collect = _timing_collect
Context.produce_range = _produce_range
Context.last_requests = _last_requests


def elev_scene(tile_w=101, grid_size=24, flip=0, noise_mode=NOISE_SLOPE, no_clamp=0, want_stats=0,
               resid_scale=1.0):
    return ElevScene(tile_w, (tile_w - 5) // grid_size, flip, noise_mode, no_clamp, want_stats,
                     resid_scale, 0)


def norm_scene(tile_w=97, grid_size=24, elev_border=2, elev_filter=FILTER_LINEAR,
               parent_filter=FILTER_LINEAR, sphere=0, arith=ARITH_EXACT):
    return NormScene(tile_w, (tile_w - 1) // grid_size, elev_border, elev_filter, parent_filter,
                     sphere, arith, 0)
