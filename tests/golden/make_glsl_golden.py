"""Writes tests/golden/glsl.json: the hashes of the tiles the reference's own GLSL shaders produce when compiled
unchanged as C++ (oracle/_ref/libref_glsl.so, built by `make -C oracle` from the shader text in /root/reference).
Run in the build container (the reference checkout must be present):  python tests/golden/make_glsl_golden.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import glsl_cases  # noqa: E402
import orc  # noqa: E402

orc.build()
assert orc.glsl() is not None, "oracle/_ref/libref_glsl.so missing: the reference checkout is needed to make the golden file"
out = {"how": "tests/golden/make_glsl_golden.py: sha1[:20] of the bytes the reference's GLSL text produces per case "
              "(fp32 elevation tiles, fp32 normal `data`, RGBA8 ortho tiles); LINEAR weights rounded to 8 fractional bits",
       "cases": glsl_cases.run(orc, glsl_cases.Engine.GLSL)}
with open(os.path.join(HERE, "glsl.json"), "w") as f:
    json.dump(out, f, indent=0, sort_keys=True)
print("%d cases" % len(out["cases"]))
