"""Writes tests/golden/ortho.json: known answers of the ORTHO oracle (oracle/orc_ortho.c).

The reference's GL path cannot run here (no GL, no Ork: SURVEY 8c), so these vectors pin the oracle
against ITSELF over time (regressions) and carry the few values that can be derived by hand from the
reference's C++ (OrthoProducer.cpp:48-118: the first interior noise byte is int(frandom(1234567) * 255)
= int(0.943745792 * 255) = 240; SURVEY 8c lists the first signed draws 2u-1 of the two border seeds,
0.05766881 and 0.30257297, i.e. u = 0.52883 and 0.65129 -> bytes 134 and 166).  Run from the repo root:
    python tests/golden/make_ortho_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import orc  # noqa: E402

TERRAIN3 = dict(hsv=1, noise_amp=[255] * 17, noise_color=[np.float32(v) / np.float32(255) for v in (70, 80, 100, 255)],
                root_noise_color=[np.float32(v) / np.float32(255) for v in (60, 150, 20, 127.5)], face=1)
PLAIN = dict(hsv=0, noise_amp=[0, 255, 255, 255, 255], noise_color=[0.5, 0, 0, 0],
             root_noise_color=[0.5, 0.5, 0.5, 0.5], face=3)


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def main():
    orc.build()
    out = {}
    for W in (196, 100):
        nz = orc.ortho_noise(W)
        out["noise_sha1_%d" % W] = [sha(nz[l]) for l in range(6)]
        out["noise_first_bytes_%d" % W] = {"interior": int(nz[0, 4, 4, 0]), "bottom_l0": int(nz[0, 2, 4, 0]),
                                          "bottom_l1": int(nz[1, 2, 4, 0])}
    for name, kw in (("terrain3_hsv", TERRAIN3), ("plain", PLAIN)):
        tiles = orc.ortho_quadtree(3, W=196, **kw)
        out[name] = {"levels_0_3_sha1": [sha(t) for t in tiles],
                     "mean_rgba": [float(v) for v in tiles.astype(np.float64).mean(axis=(0, 1, 2))]}
    with open(os.path.join(HERE, "ortho.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote ortho.json")


if __name__ == "__main__":
    main()
