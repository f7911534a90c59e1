"""ctypes face of the CPU oracle (oracle/liborc.so) and of the reference's own
noise.cpp (oracle/_ref/libref_noise.so).  TEST INFRASTRUCTURE: imported only by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORC_DIR = os.path.join(ROOT, "oracle")
ORC_SO = os.path.join(ORC_DIR, "liborc.so")
REF_SO = os.path.join(ORC_DIR, "_ref", "libref_noise.so")
GLSL_SO = os.path.join(ORC_DIR, "_ref", "libref_glsl.so")
HM_SO = os.path.join(ORC_DIR, "_ref", "libref_hm.so")
STRICT_SO = os.path.join(ORC_DIR, "liborc_strict.so")

c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int)
c_u8_p = C.POINTER(C.c_uint8)


def build():
    """(Re)build the oracle; also rebuilds oracle/_ref when /root/reference exists."""
    subprocess.check_call(["make", "-s", "-C", ORC_DIR], stdout=subprocess.DEVNULL)


class ElevParams(C.Structure):
    _fields_ = [("W", C.c_int), ("level", C.c_int), ("pixel_size", C.c_float),
                ("grid", C.c_int), ("flip", C.c_int), ("dx", C.c_int), ("dy", C.c_int),
                ("has_resid", C.c_int), ("rx", C.c_int), ("ry", C.c_int),
                ("resid_stride", C.c_int), ("noiseR", C.c_int), ("noiseL", C.c_int),
                ("rs", C.c_float), ("noise_mode", C.c_int), ("no_clamp", C.c_int)]


class NormParams(C.Structure):
    _fields_ = [("W", C.c_int), ("grid", C.c_int), ("format", C.c_int),
                ("elev_W", C.c_int), ("elev_border", C.c_int), ("elev_filter", C.c_int),
                ("has_parent", C.c_int), ("ptx", C.c_int), ("pty", C.c_int),
                ("parent_filter", C.c_int),
                ("deform", C.c_float * 4), ("corners", C.c_float * 16),
                ("verticals", C.c_float * 16), ("norms", C.c_float * 4),
                ("w2t", C.c_float * 9), ("p2t", C.c_float * 9)]


class ResidFile(C.Structure):
    pass


ResidFile._fields_ = [("minLevel", C.c_int), ("maxLevel", C.c_int), ("tileSize", C.c_int),
                      ("rootLevel", C.c_int), ("rootTx", C.c_int), ("rootTy", C.c_int),
                      ("scale", C.c_float), ("deltaLevel", C.c_int), ("ntiles", C.c_int),
                      ("header", C.c_uint32), ("offsets", C.c_void_p), ("data", C.c_void_p),
                      ("size", C.c_size_t), ("nchildren", C.c_int),
                      ("children", C.POINTER(C.POINTER(ResidFile)))]


class Scene(C.Structure):
    _fields_ = [("W", C.c_int), ("gridMeshSize", C.c_int), ("rootQuadSize", C.c_float),
                ("face", C.c_int), ("flip", C.c_int), ("noise_mode", C.c_int),
                ("no_clamp", C.c_int), ("nAmp", C.c_int), ("noiseAmp", C.c_float * 32),
                ("sphere", C.c_int), ("elev_filter", C.c_int),
                ("resid", C.POINTER(ResidFile))]


class OrthoParams(C.Structure):
    _fields_ = [("tileWidth", C.c_int), ("level", C.c_int), ("dx", C.c_int), ("dy", C.c_int),
                ("hasResidual", C.c_int), ("residualScale", C.c_float), ("noiseR", C.c_int), ("noiseL", C.c_int),
                ("hsv", C.c_int), ("noiseColor", C.c_float * 4), ("rootNoiseColor", C.c_float * 4)]


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORC_SO):
            build()
        L = C.CDLL(ORC_SO)
        L.orc_lrandom.restype = C.c_long
        L.orc_lrandom.argtypes = [C.POINTER(C.c_long)]
        L.orc_frandom.restype = C.c_float
        L.orc_frandom.argtypes = [C.POINTER(C.c_long)]
        L.orc_cnoise2.restype = C.c_float
        L.orc_cnoise2.argtypes = [C.c_float, C.c_float]
        L.orc_round_half.restype = C.c_float
        L.orc_round_half.argtypes = [C.c_float]
        L.orc_float_to_half_bits.restype = C.c_uint16
        L.orc_float_to_half_bits.argtypes = [C.c_float]
        L.orc_unorm8.restype = C.c_uint8
        L.orc_unorm8.argtypes = [C.c_float]
        L.orc_tiff_inflate.restype = C.c_long
        L.orc_resid_blob.restype = C.c_void_p
        L.orc_produce_quadtree.restype = C.c_long
        _lib = L
    return _lib


def ref():
    """The reference's own noise.cpp, or None when oracle/_ref was not built."""
    global _ref
    if _ref is None:
        if not os.path.exists(REF_SO):
            return None
        R = C.CDLL(REF_SO)
        R.ref_lrandom.restype = C.c_long
        R.ref_lrandom.argtypes = [C.POINTER(C.c_long)]
        R.ref_frandom.restype = C.c_float
        R.ref_frandom.argtypes = [C.POINTER(C.c_long)]
        R.ref_cnoise2.restype = C.c_float
        R.ref_cnoise2.argtypes = [C.c_float, C.c_float]
        _ref = R
    return _ref


_strict = None
_glsl = None


def strict():
    """The non-contracted reading of the restatement (liborc_strict.so): every a*b+c is two roundings.  It is
    the reading that must equal the reference's shader text compiled as plain IEEE fp32 (glsl()) bit for bit."""
    global _strict
    if _strict is None:
        if not os.path.exists(STRICT_SO):
            build()
        _strict = C.CDLL(STRICT_SO)
    return _strict


def glsl():
    """The reference's own GLSL shaders compiled unchanged as C++ (oracle/_ref/libref_glsl.so, oracle/Makefile),
    or None when oracle/_ref was not built."""
    global _glsl
    if _glsl is None:
        if not os.path.exists(GLSL_SO):
            return None
        _glsl = C.CDLL(GLSL_SO)
    return _glsl


_hm = None


def hm():
    """The reference's own residual builder (preprocess/terrain/*.cpp compiled unchanged over the Ork / libtiff shims:
    oracle/_ref/libref_hm.so, oracle/Makefile), or None when oracle/_ref was not built."""
    global _hm
    if _hm is None:
        if not os.path.exists(HM_SO):
            return None
        H = C.CDLL(HM_SO)
        for f in (H.ref_preprocess_spherical_dem, H.ref_preprocess_dem):
            f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_float]
        _hm = H
    return _hm


def ref_preprocess_dem(src, min_tile_size, tile_size, max_level, dst, tmp, scale=1.0, spherical=False):
    """proland::preprocessDem / preprocessSphericalDem of the reference itself -> dst/DEM.dat or dst/DEM1..6.dat"""
    src = np.ascontiguousarray(src, np.float32)
    f = hm().ref_preprocess_spherical_dem if spherical else hm().ref_preprocess_dem
    rc = f(src.ctypes.data, src.shape[1], src.shape[0], min_tile_size, tile_size, max_level, os.fsencode(dst), os.fsencode(tmp), scale)
    if rc != 0:
        raise RuntimeError("the reference's preprocess failed")


def _fp(a):
    return a.ctypes.data_as(c_float_p)


def _fpn(a):
    return _fp(np.ascontiguousarray(a, np.float32)) if a is not None else None


UPSAMPLE_VARIANT = {"A": 0, "B": 1, "C": 2, "D": 3, "D_NO_CLAMP": 4}
NORMAL_VARIANT = {"flat": 0, "sphere": 1, "demo": 2}


def glsl_upsample_tile(variant, p, parent, resid, noise, parent_filter=1, subtexel_bits=8):
    """upsampleShader.glsl (the reference's text) on one tile -> (W, W, 3) float32"""
    out = np.empty((p.W, p.W, 3), np.float32)
    par = np.ascontiguousarray(parent, np.float32) if parent is not None else None
    res = np.ascontiguousarray(resid, np.float32) if resid is not None else None
    rc = glsl().ref_glsl_upsample_tile(C.c_int(UPSAMPLE_VARIANT[variant]), C.byref(p), _fpn(par), C.c_int(parent_filter),
                                       _fpn(res), _fp(noise), C.c_int(subtexel_bits), _fp(out))
    assert rc == 0
    return out


def glsl_normal_tile(variant, p, elev, parent=None, subtexel_bits=8):
    """normalShader.glsl (the reference's text) on one tile -> (W, W, 4) float32, the fragment's `data`"""
    out = np.empty((p.W, p.W, 4), np.float32)
    el = np.ascontiguousarray(elev, np.float32)
    par = np.ascontiguousarray(parent, np.float32) if parent is not None else None
    rc = glsl().ref_glsl_normal_tile(C.c_int(NORMAL_VARIANT[variant]), C.byref(p), _fp(el), _fpn(par),
                                     C.c_int(subtexel_bits), _fp(out))
    assert rc == 0
    return out


def glsl_ortho_tile(variant, p, parent, residual, noise, channels=4, parent_filter=0):
    """upsampleOrthoShader.glsl (the reference's text) on one tile -> (W, W, 4) float32 `data`"""
    W = p.tileWidth
    out = np.empty((W, W, 4), np.float32)
    par = np.ascontiguousarray(parent, np.uint8) if parent is not None else None
    res = np.ascontiguousarray(residual, np.uint8) if residual is not None else None
    nz = np.ascontiguousarray(noise, np.uint8)
    rc = glsl().ref_glsl_ortho_tile(C.c_int(variant), C.byref(p), _u8(par), C.c_int(parent_filter), _u8(res),
                                    C.c_int(channels), _u8(nz), _fp(out))
    assert rc == 0
    return out


# ---------------------------------------------------------------- noise ----

def dem_noise(W=101, r16f=True):
    out = np.empty((6, W, W), np.float32)
    (lib().orc_dem_noise_r16f if r16f else lib().orc_dem_noise)(C.c_int(W), _fp(out))
    return out


def cnoise_tables():
    p = np.empty(514, np.int32)
    g2 = np.empty((514, 2), np.float32)
    lib().orc_cnoise_tables(p.ctypes.data_as(c_int_p), _fp(g2))
    return p, g2


def noise_select(level, tx, ty, face):
    r, l = C.c_int(), C.c_int()
    lib().orc_noise_select(level, tx, ty, face, C.byref(r), C.byref(l))
    return r.value, l.value


# ------------------------------------------------------------ elevation ----

def elev_uniforms(level, tx, ty, *, W=101, gridMeshSize=24, rootQuadSize=100000.0, flip=0,
                  noiseAmp=(), face=0, has_resid=0, resid_W=0, noise_mode=1, no_clamp=0):
    p = ElevParams()
    amp = np.asarray(noiseAmp, np.float32)
    lib().orc_elev_uniforms(W, gridMeshSize, C.c_float(rootQuadSize), flip, _fp(amp), len(amp),
                            face, level, tx, ty, has_resid, resid_W, noise_mode, no_clamp,
                            C.byref(p))
    return p


def upsample_tile(p, parent, resid, noise, L=None):
    W = p.W
    out = np.empty((W, W, 3), np.float32)
    par = np.ascontiguousarray(parent, np.float32) if parent is not None else None
    res = np.ascontiguousarray(resid, np.float32) if resid is not None else None
    (L or lib()).orc_upsample_tile(C.byref(p), _fpn(par), _fpn(res), _fp(noise), _fp(out))
    return out


def cpu_elevation_tile(W, level, tx, ty, parent, resid, resid_W, rx, ry):
    out = np.empty((W, W), np.float32)
    par = _fp(np.ascontiguousarray(parent, np.float32)) if parent is not None else None
    res = _fp(np.ascontiguousarray(resid, np.float32)) if resid is not None else None
    lib().orc_cpu_elevation_tile(W, level, tx, ty, par, res, resid_W, rx, ry, _fp(out))
    return out


def tile_minmax(elev):
    W = elev.shape[0]
    a, b = C.c_float(), C.c_float()
    lib().orc_tile_minmax(W, _fp(np.ascontiguousarray(elev, np.float32)), C.byref(a), C.byref(b))
    return a.value, b.value


# -------------------------------------------------------------- normals ----

def normal_uniforms(level, tx, ty, *, W=97, gridMeshSize=24, components=2, signed_comp=0,
                    elev_W=101, elev_border=2, elev_filter=1, parent_filter=1,
                    rootQuadSize=100000.0, sphere=0):
    p = NormParams()
    lib().orc_normal_uniforms(W, gridMeshSize, components, signed_comp, elev_W, elev_border,
                              elev_filter, parent_filter, C.c_double(rootQuadSize), sphere,
                              level, tx, ty, C.byref(p))
    return p


def normal_tile(p, elev, parent=None, L=None):
    out = np.empty((p.W, p.W, 4), np.float32)
    par = np.ascontiguousarray(parent, np.float32) if parent is not None else None
    el = np.ascontiguousarray(elev, np.float32)
    (L or lib()).orc_normal_tile(C.byref(p), _fp(el), _fpn(par), _fp(out))
    return out


def pack_unorm8(data, channels):
    W = data.shape[0]
    out = np.empty((W, W, channels), np.uint8)
    lib().orc_pack_unorm8(W, channels, _fp(np.ascontiguousarray(data, np.float32)),
                          out.ctypes.data_as(c_u8_p))
    return out


# ------------------------------------------------------------ residuals ----

class Resid:
    """An opened residual container (keeps the file bytes alive)."""

    def __init__(self, data: bytes, delta=0, zscale=1.0):
        self.buf = np.frombuffer(data, np.uint8).copy()
        self.f = ResidFile()
        rc = lib().orc_resid_open(self.buf.ctypes.data_as(c_u8_p), C.c_size_t(len(self.buf)),
                                  delta, C.c_float(zscale), C.byref(self.f))
        if rc != 0:
            raise ValueError("bad residual container rc=%d" % rc)
        self._children = []

    def add_child(self, child):
        self._children.append(child)
        arr = (C.POINTER(ResidFile) * len(self._children))(
            *[C.pointer(c.f) for c in self._children])
        self._arr = arr
        self.f.children = C.cast(arr, C.POINTER(C.POINTER(ResidFile)))
        self.f.nchildren = len(self._children)

    @property
    def width(self):
        return self.f.tileSize + 5

    def has_tile(self, level, tx, ty):
        return bool(lib().orc_resid_has_tile(C.byref(self.f), level, tx, ty))

    def tile_id(self, l, tx, ty):
        return lib().orc_resid_tile_id(C.byref(self.f), l, tx, ty)

    def tile_size(self, l):
        return lib().orc_resid_tile_size(C.byref(self.f), l)

    def blob(self, tileid):
        n = C.c_uint32()
        p = lib().orc_resid_blob(C.byref(self.f), tileid, C.byref(n))
        off = p - self.buf.ctypes.data
        return self.buf[off:off + n.value].tobytes()

    def inflate(self, tileid):
        b = np.frombuffer(self.blob(tileid), np.uint8)
        cap = self.width * self.width * 2
        raw = np.empty(cap, np.uint8)
        w, h = C.c_int(), C.c_int()
        n = lib().orc_tiff_inflate(b.ctypes.data_as(c_u8_p), C.c_uint32(len(b)),
                                   raw.ctypes.data_as(c_u8_p), C.c_size_t(cap),
                                   C.byref(w), C.byref(h))
        if n < 0:
            raise ValueError("inflate failed rc=%d" % n)
        return raw[:n].tobytes(), w.value, h.value

    def create_tile(self, level, tx, ty):
        n = self.width
        out = np.zeros((n, n), np.float32)
        rc = lib().orc_resid_create_tile(C.byref(self.f), level, tx, ty, _fp(out))
        if rc != 0:
            raise ValueError("create_tile rc=%d" % rc)
        return out


def hm_encode_tile(parent, tile, tile_size, tx, ty):
    """HeightMipmap computeResidual + encodeResidual + computeApproxTile for one tile (orc_preprocess.c).
    parent, tile: (n, n) float32 with n = the container's tileSize + 5.
    -> (int16 residuals (ts+5, ts+5), approximation (n, n), max |residual|, max |tile - approximation|)"""
    parent = np.ascontiguousarray(parent, np.float32)
    tile = np.ascontiguousarray(tile, np.float32)
    n = tile.shape[0]
    w = tile_size + 5
    resid = np.zeros((w, w), np.int16)
    approx = np.zeros((n, n), np.float32)
    mr, me = C.c_float(), C.c_float()
    L = lib()
    L.orc_hm_encode_tile.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_hm_encode_tile(parent.ctypes.data, tile.ctypes.data, n, tile_size, tx, ty, resid.ctypes.data,
                         approx.ctypes.data, C.byref(mr), C.byref(me))
    return resid, approx, mr.value, me.value


def _face_ptrs(faces):
    faces = [np.ascontiguousarray(f, np.int16) for f in faces]
    return faces, (C.c_void_p * len(faces))(*[f.ctypes.data for f in faces])


def hm_height(faces, max_level, level, face, x, y):
    """HeightMipmap::getTileHeight with setCube's stitching (orc_preprocess.c); faces: 1 or 6 (B + 1, B + 1) int16 grids"""
    faces, ptrs = _face_ptrs(faces)
    L = lib()
    L.orc_hm_height.restype = C.c_float
    L.orc_hm_height.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    return L.orc_hm_height(ptrs, len(faces), faces[0].shape[0] - 1, max_level, level, face, x, y)


def hm_get_tile(faces, max_level, top_level_size, tile_size, level, face, tx, ty, scale=1.0):
    """HeightMipmap::getTile -> (tileSize + 5, tileSize + 5) float32, the tile in the lower-left (ts + 5)^2"""
    faces, ptrs = _face_ptrs(faces)
    n = tile_size + 5
    out = np.zeros((n, n), np.float32)
    L = lib()
    L.orc_hm_get_tile.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int,
                                  C.c_int, C.c_int, C.c_void_p]
    L.orc_hm_get_tile(ptrs, len(faces), faces[0].shape[0] - 1, max_level, top_level_size, tile_size, scale, level, face, tx, ty,
                      out.ctypes.data)
    return out


def cube_projection(face, x, y, w):
    """projection1..6 of Preprocess.cpp:155-213 -> (sx, sy, sz)"""
    L = lib()
    L.orc_cube_projection.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    v = (C.c_double * 3)()
    L.orc_cube_projection(face, x, y, w, C.byref(v, 0), C.byref(v, 8), C.byref(v, 16))
    return tuple(v)


def spherical_base(src, face, B):
    """the base grid of one cube face from an equirectangular map: (short) SphericalHeightFunction::getHeight(x, y)"""
    src = np.ascontiguousarray(src, np.float32)
    L = lib()
    L.orc_spherical_base_grid.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    out = np.empty((B + 1, B + 1), np.int16)
    L.orc_spherical_base_grid(src.ctypes.data, src.shape[1], src.shape[0], face, B, out.ctypes.data)
    return out


def plane_base(src, B):
    """the base grid of a flat DEM: (short) PlaneHeightFunction::getHeight(x, y)"""
    src = np.ascontiguousarray(src, np.float32)
    L = lib()
    L.orc_plane_base_grid.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    out = np.empty((B + 1, B + 1), np.int16)
    L.orc_plane_base_grid(src.ctypes.data, src.shape[1], src.shape[0], B, out.ctypes.data)
    return out


# --------------------------------------------------------------- driver ----

def make_scene(*, W=101, gridMeshSize=24, rootQuadSize=100000.0, face=0, flip=0, noise_mode=1,
               no_clamp=0, noiseAmp=(), sphere=0, elev_filter=1, resid=None):
    s = Scene()
    s.W, s.gridMeshSize, s.rootQuadSize, s.face = W, gridMeshSize, rootQuadSize, face
    s.flip, s.noise_mode, s.no_clamp = flip, noise_mode, no_clamp
    s.nAmp = len(noiseAmp)
    for i, a in enumerate(noiseAmp):
        s.noiseAmp[i] = a
    s.sphere, s.elev_filter = sphere, elev_filter
    s.resid = C.pointer(resid.f) if resid is not None else None
    s._keep = resid
    return s


def produce_pair(scene, noise, level, tx, ty, parent, resid_tile=None, want_normals=True):
    W = scene.W
    elev = np.empty((W, W, 3), np.float32)
    norm = np.empty((W - 4, W - 4, 2), np.uint8) if want_normals else None
    par = _fp(np.ascontiguousarray(parent, np.float32)) if parent is not None else None
    res = _fp(np.ascontiguousarray(resid_tile, np.float32)) if resid_tile is not None else None
    lib().orc_produce_pair(C.byref(scene), _fp(noise), level, tx, ty, par, res, _fp(elev),
                           norm.ctypes.data_as(c_u8_p) if want_normals else None)
    return elev, norm


def produce_quadtree(scene, max_level, nthreads=0):
    cs, lo, hi = C.c_double(), C.c_float(), C.c_float()
    n = lib().orc_produce_quadtree(C.byref(scene), max_level, nthreads, C.byref(cs),
                                   C.byref(lo), C.byref(hi))
    return n, cs.value, lo.value, hi.value


def glsl_produce_quadtree(scene, max_level, nthreads=0):
    """the same quadtree by the reference's own shader text (oracle/_ref/libref_glsl.so, ref_glsl_produce_quadtree);
    None when oracle/_ref was not built"""
    G = glsl()
    if G is None or not hasattr(G, "ref_glsl_produce_quadtree"):
        return None
    G.ref_glsl_produce_quadtree.restype = C.c_long
    cs, lo, hi = C.c_double(), C.c_float(), C.c_float()
    n = G.ref_glsl_produce_quadtree(C.byref(scene), max_level, nthreads, C.byref(cs), C.byref(lo), C.byref(hi))
    return (n, cs.value, lo.value, hi.value) if n > 0 else None


# ---------------------------------------------------------------- ortho ----

def _u8(a):
    return a.ctypes.data_as(c_u8_p) if a is not None else None


def _f4(v):
    return (C.c_float * 4)(*[float(x) for x in v])


def ortho_noise(W=196):
    out = np.empty((6, W, W, 4), np.uint8)
    lib().orc_ortho_noise(C.c_int(W), _u8(out))
    return out


def ortho_uniforms(level, tx, ty, *, W=196, face=1, noise_amp=(), noise_color=(1, 1, 1, 1),
                   root_noise_color=(0.5, 0.5, 0.5, 0.5), hsv=0, scale=2.0, has_residual=0):
    p = OrthoParams()
    amp = np.asarray(noise_amp, np.float32)
    lib().orc_ortho_uniforms(C.c_int(W), C.c_int(face), C.c_int(level), C.c_int(tx), C.c_int(ty), _fp(amp),
                             C.c_int(len(amp)), _f4(noise_color), _f4(root_noise_color), C.c_int(int(hsv)),
                             C.c_float(scale), C.c_int(int(has_residual)), C.byref(p))
    return p


def ortho_tile(p, parent, residual, noise, channels=4, L=None):
    """parent: (W, W, 4) uint8 or None; residual: (W, W, channels) uint8 or None -> (W, W, 4) uint8"""
    W = p.tileWidth
    out = np.empty((W, W, 4), np.uint8)
    par = np.ascontiguousarray(parent, np.uint8) if parent is not None else None
    res = np.ascontiguousarray(residual, np.uint8) if residual is not None else None
    nz = np.ascontiguousarray(noise, np.uint8)
    (L or lib()).orc_ortho_tile(C.byref(p), _u8(par), _u8(res), C.c_int(channels), _u8(nz), _u8(out))
    return out


def ortho_quadtree(max_level, *, W=196, face=1, noise_amp=(), noise_color=(1, 1, 1, 1),
                   root_noise_color=(0.5, 0.5, 0.5, 0.5), hsv=0, scale=2.0, noise=None):
    """All tiles of levels 0..max_level, level order, Morton order inside a level -> (n, W, W, 4) uint8"""
    n = (4 ** (max_level + 1) - 1) // 3
    out = np.empty((n, W, W, 4), np.uint8)
    amp = np.asarray(noise_amp, np.float32)
    nz = np.ascontiguousarray(noise if noise is not None else ortho_noise(W), np.uint8)
    L = lib()
    L.orc_ortho_quadtree.restype = C.c_long
    done = L.orc_ortho_quadtree(C.c_int(W), C.c_int(face), C.c_int(max_level), _fp(amp), C.c_int(len(amp)),
                                _f4(noise_color), _f4(root_noise_color), C.c_int(int(hsv)), C.c_float(scale),
                                _u8(nz), _u8(out))
    assert done == n
    return out


def ortho_cpu_read(file_bytes, level, tx, ty, W=196, max_channels=4):
    """OrthoCPUProducer's reader on a whole file image -> (W, W, channels) uint8, or the negative error code"""
    buf = np.frombuffer(file_bytes, np.uint8)
    out = np.empty(W * W * max_channels, np.uint8)
    ch = C.c_int(0)
    rc = lib().orc_ortho_cpu_read(_u8(buf), C.c_size_t(len(file_bytes)), C.c_int(level), C.c_int(tx), C.c_int(ty),
                                  _u8(out), C.byref(ch))
    if rc < 0:
        return rc
    return out[:rc * rc * ch.value].reshape(rc, rc, ch.value).copy()
