"""Writes tests/golden/hm.json: per case of tests/hm_cases.py, the digest of the residual files the reference's own
builder writes (preprocess/terrain/*.cpp compiled unchanged: oracle/_ref/libref_hm.so, `make -C oracle`).
Run in the build container (the reference checkout must be present):  python tests/golden/make_hm_golden.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import hm_cases  # noqa: E402
import orc  # noqa: E402

orc.build()
assert orc.hm() is not None, "oracle/_ref/libref_hm.so missing: the reference checkout is needed to make the golden file"
out = {"how": "tests/golden/make_hm_golden.py: sha1[:20] per face over (header, blob sharing, int16 tiles in id order) of the "
              "files proland::preprocessDem / preprocessSphericalDem of the reference write for each case",
       "cases": {name: hm_cases.digest(hm_cases.reference_record(orc, name)) for name in hm_cases.CASES}}
with open(os.path.join(HERE, "hm.json"), "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)
print(json.dumps(out["cases"], indent=1))
