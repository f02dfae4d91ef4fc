#!/usr/bin/env python
"""Build a variant of libb200smpl.so with extra nvcc flags for some sources (experiments):
   build_variant.py <name> <source.cu[,source.cu...]> <flags...>   ->  _variants/<name>/libb200smpl.so
   (run with B200SMPL_LIB=_variants/<name>/libb200smpl.so)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from soccerplayershapepose_b200 import build as B    # noqa: E402
name, srcs, flags = sys.argv[1], sys.argv[2].split(","), sys.argv[3:]
B.build_library()
out = os.path.join(ROOT, "_variants", name)
os.makedirs(out, exist_ok=True)
objs = []
for src in B.SOURCES:
    o = os.path.join(B.BUILD, src.replace(".cu", ".o"))
    if src in srcs:
        o = os.path.join(out, src.replace(".cu", ".o"))
        cmd = [B._nvcc()] + B.NVCC_FLAGS + flags + ["-c", os.path.join(B.CSRC, src), "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        open(o + ".log", "w").write(r.stdout + r.stderr)
        if r.returncode != 0:
            sys.exit(r.stdout + r.stderr)
        print("\n".join(l for l in (r.stdout + r.stderr).splitlines() if "Used" in l or "spill" in l))
    objs.append(o)
lib = os.path.join(out, "libb200smpl.so")
subprocess.check_call([B._nvcc(), "-shared", "-o", lib] + objs + ["-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a"])
print(lib)
