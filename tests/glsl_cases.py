"""The cases that pin the float part of the oracle on the reference's own shader text.  TEST INFRASTRUCTURE.

run(engine) produces every case with one engine and returns {case name: sha1 of the produced bytes}:
  Engine.STRICT  oracle/liborc_strict.so  -- the restatement, no contraction (one rounding per GLSL operation)
  Engine.GLSL    oracle/_ref/libref_glsl.so -- upsampleShader.glsl / normalShader.glsl / upsampleOrthoShader.glsl
                 of the reference checkout compiled unchanged as C++ (oracle/ref_glsl_wrap.cpp)
The two must agree bit for bit (tests/test_glsl_pin.py); tests/golden/glsl.json holds the GLSL engine's hashes
(tests/golden/make_glsl_golden.py) so that the pin also holds where the reference checkout is absent.

Chains run root -> leaf, every engine feeding on its own parents."""
import hashlib

import numpy as np

FRACTAL = [-140, -100, -15, -8, 5, 2.5, 1.5, 1, 0.5, 0.25, 0.1, 0.05]
PLANET = [-3250, -1590, -1125, -795, -561, -397, -140, -100, 15, 8, 5, 2.5, 1.5, 1, 0.5, 0.25, 0.1, 0.05]
SRTM = [0] * 11 + [5, 2.5, 1, 0.5, 0.25, 0.1, 0.05, 0.025, 0.01, 0.01, 0.005, 0.005]
NEAREST, LINEAR = 0, 1


def chain(level, tx, ty):
    """the root-to-(level, tx, ty) chain of tiles"""
    return [(l, tx >> (level - l), ty >> (level - l)) for l in range(level + 1)]


def _sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()[:20]


class Engine:
    STRICT, GLSL = "strict", "glsl"

    def __init__(self, orc, kind):
        self.orc, self.kind = orc, kind

    def up(self, variant, p, parent, resid, noise, parent_filter):
        if self.kind == Engine.GLSL:
            return self.orc.glsl_upsample_tile(variant, p, parent, resid, noise, parent_filter)
        return self.orc.upsample_tile(p, parent, resid, noise, L=self.orc.strict())

    def nrm(self, variant, p, elev, parent=None):
        if self.kind == Engine.GLSL:
            return self.orc.glsl_normal_tile(variant, p, elev, parent)
        return self.orc.normal_tile(p, elev, parent, L=self.orc.strict())

    def ortho(self, variant, p, parent, resid, noise, channels, parent_filter):
        if self.kind == Engine.GLSL:
            data = self.orc.glsl_ortho_tile(variant, p, parent, resid, noise, channels, parent_filter)
            return self.orc.pack_unorm8(data, 4)      # the RGBA8 colour buffer
        return self.orc.ortho_tile(p, parent, resid, noise, channels, L=self.orc.strict())


# variant letter -> (noise_mode, flip honoured, no_clamp) of the restatement
VARIANTS = {"A": (0, 0, 0), "B": (1, 0, 0), "C": (0, 1, 0), "D": (1, 1, 0), "D_NO_CLAMP": (1, 1, 1)}


def elevation_cases(eng, out, full=True):
    orc = eng.orc
    noise = orc.dem_noise(101)
    rng = np.random.default_rng(20240612)
    resid = np.round(rng.normal(0, 40, (197, 197))).astype(np.float32)
    # (name, variant, amplitudes, face, root quad size, flip attribute, leaf, elevation storage filter, residuals)
    runs = [
        ("fractalterrain", "D", FRACTAL, 0, 100000.0, 0, (8, 201, 77), LINEAR, False),          # config 1
        ("fractalplanet_f3", "D", PLANET, 3, 12720000.0, 0, (10, 750, 413), LINEAR, False),     # config 2
        ("fractalplanet_f1", "D", PLANET, 1, 12720000.0, 0, (10, 1023, 0), LINEAR, False),
        ("fractalplanet_f6", "D", PLANET, 6, 12720000.0, 0, (9, 0, 511), LINEAR, False),
        ("earth_srtm_f2", "D", SRTM, 2, 12720000.0, 1, (12, 2901, 1717), NEAREST, True),        # config 3
        ("subtree_l14", "D", FRACTAL + [0, 0, 0], 0, 100000.0, 0, (14, 9999, 12345), LINEAR, False),   # config 4
        ("variant_a", "A", FRACTAL[:2] + [30, 20, 10], 0, 100000.0, 0, (4, 9, 6), LINEAR, True),
        ("variant_b", "B", FRACTAL[:2] + [30, 20, 10], 0, 100000.0, 0, (4, 6, 9), NEAREST, False),
        ("variant_c", "C", FRACTAL[:2] + [30, 20, 10], 0, 100000.0, 1, (4, 3, 12), LINEAR, True),
        ("variant_d_no_clamp", "D_NO_CLAMP", FRACTAL[:2] + [30, 20, 10], 0, 100000.0, 1, (4, 12, 3), LINEAR, False),
    ]
    if not full:
        runs = [r for r in runs if r[0] in ("fractalplanet_f3", "earth_srtm_f2", "variant_c")]
    tiles = {}
    for name, variant, amp, face, rqs, flip, leaf, filt, with_resid in runs:
        mode, honours_flip, no_clamp = VARIANTS[variant]
        parent = None
        for (l, tx, ty) in chain(*leaf):
            has_resid = int(with_resid and l >= 1)
            p = orc.elev_uniforms(l, tx, ty, rootQuadSize=rqs, noiseAmp=amp, face=face, flip=int(flip and honours_flip),
                                  noise_mode=mode, no_clamp=no_clamp, has_resid=has_resid, resid_W=197 if has_resid else 0)
            e = eng.up(variant, p, parent, resid if has_resid else None, noise, filt)
            # the coarse height of the first two rows / columns reads the parent at index -2 (clamped): dead texels
            # of the reference too, kept in the hash because both engines clamp to the edge
            out["elev/%s/%d_%d_%d" % (name, l, tx, ty)] = _sha(e)
            tiles[(name, l)] = (e, rqs, face, filt)
            parent = e
    return tiles


def normal_cases(eng, out, tiles, full=True):
    orc = eng.orc
    for (name, l), (e, rqs, face, filt) in sorted(tiles.items()):
        if not full and l % 3:
            continue
        sphere = int(face != 0)
        leafs = {"fractalterrain": (8, 201, 77), "fractalplanet_f3": (10, 750, 413), "fractalplanet_f1": (10, 1023, 0),
                 "fractalplanet_f6": (9, 0, 511), "earth_srtm_f2": (12, 2901, 1717), "subtree_l14": (14, 9999, 12345)}
        if name not in leafs:
            continue
        L, TX, TY = leafs[name]
        tx, ty = TX >> (L - l), TY >> (L - l)
        q = orc.normal_uniforms(l, tx, ty, rootQuadSize=rqs, sphere=sphere, elev_filter=filt)
        for variant in (("demo", "sphere") if sphere else ("demo", "flat")):
            out["norm/%s/%s/%d" % (name, variant, l)] = _sha(eng.nrm(variant, q, e))
    # the four output formats and the parent coarse normal (normalShader.glsl:100-124), demo shader only
    rng = np.random.default_rng(7)
    for (name, l) in (("fractalplanet_f3", 5), ("fractalterrain", 4)):
        if (name, l) not in tiles:
            continue
        e, rqs, face, filt = tiles[(name, l)]
        sphere = int(face != 0)
        L, TX, TY = (10, 750, 413) if sphere else (8, 201, 77)
        tx, ty = TX >> (L - l), TY >> (L - l)
        for components, signed in ((4, 1), (4, 0), (2, 1), (2, 0)):
            for pfilt in (NEAREST, LINEAR):
                q = orc.normal_uniforms(l, tx, ty, rootQuadSize=rqs, sphere=sphere, elev_filter=filt, components=components,
                                        signed_comp=signed, parent_filter=pfilt)
                # a parent normal tile as its sampler returns it: unit-ish xy in [-0.6, 0.6] (signed) or [0.2, 0.8]
                pn = rng.uniform(-0.6, 0.6, (97, 97, 4)).astype(np.float32)
                if not signed:
                    pn = (np.round((pn * 0.5 + 0.5) * 255) / 255).astype(np.float32)
                out["norm/%s/fmt%d%d_pf%d" % (name, components, signed, pfilt)] = _sha(eng.nrm("demo", q, e, pn if components == 4 else None))


def ortho_cases(eng, out, full=True):
    orc = eng.orc
    for W, max_level in ((196, 4), (100, 2)):
        noise = orc.ortho_noise(W)
        for hsv in (1, 0):
            kw = dict(W=W, face=1 if hsv else 4, noise_amp=[255, 200, 160, 120, 90],
                      noise_color=[np.float32(v) / np.float32(255) for v in ((70, 80, 100, 255) if hsv else (255, 255, 255, 255))],
                      root_noise_color=[np.float32(v) / np.float32(255) for v in (60, 150, 20, 127.5)], hsv=hsv, scale=2.0)
            parent = None
            for (l, tx, ty) in chain(max_level, 11 >> (4 - max_level), 6 >> (4 - max_level)):
                nxt = None
                for channels in (0, 4, 3, 1):      # 0: no residual tile
                    if not full and channels in (3, 1):
                        continue
                    p = orc.ortho_uniforms(l, tx, ty, has_residual=int(channels > 0), **kw)
                    rng = np.random.default_rng([5, W, hsv, l, channels])
                    res = rng.integers(96, 160, (W, W, channels), dtype=np.uint8) if channels else None
                    for variant in (0, 1):
                        for pfilt in (NEAREST, LINEAR):
                            t = eng.ortho(variant, p, parent, res, noise, max(channels, 1), pfilt)
                            out["ortho/%d/hsv%d/%d_%d_%d/ch%d/v%d_pf%d" % (W, hsv, l, tx, ty, channels, variant, pfilt)] = _sha(t)
                    if channels == 0:
                        nxt = t
                parent = nxt


def run(orc, kind, full=True):
    eng = Engine(orc, kind)
    out = {}
    tiles = elevation_cases(eng, out, full)
    normal_cases(eng, out, tiles, full)
    ortho_cases(eng, out, full)
    return out
