#!/usr/bin/env python
"""Build experiment variants of libhode.so (extra -D flags on hode_rollout_tc.cu) into build/variants/ so that one
GPU call can time several of them: tools/tc_mode_errors.py picks one with HODE_LIB_PATH.
Usage: python tools/build_variants.py name1:-DFLAG_A,-DFLAG_B name2: ..."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybrid_ode_for_glp_1_and_glucose_b200 import build as B
B.build()
out = os.path.join(ROOT, "build", "variants")
os.makedirs(out, exist_ok=True)
for spec in sys.argv[1:]:
    name, _, flags = spec.partition(":")
    flags = [f for f in flags.split(",") if f]
    obj = os.path.join(out, f"hode_rollout_tc_{name}.o")
    subprocess.run([B._nvcc(), *B.NVCC_FLAGS, *flags, "-c", os.path.join(B.CSRC, "hode_rollout_tc.cu"), "-o", obj], check=True)
    objs = [os.path.join(B.CSRC, s.replace(".cu", ".o")) for s in B.SOURCES if s != "hode_rollout_tc.cu"] + [obj]
    lib = os.path.join(out, f"libhode_{name}.so")
    subprocess.run([B._nvcc(), "-shared", "-o", lib, *objs], check=True)
    print(lib)
