#!/usr/bin/env python3
"""Kernel tuning aid: build the library with extra -D flags into build_exp/lib_<name>.so (loaded with
SMART_B200_LIB=...).  Usage: python tools/build_variant.py <name> [-DFOO=1 ...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from smartpy_b200 import _build  # noqa: E402

name, extra = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(ROOT, "build_exp")
obj_dir = os.path.join(out_dir, "obj_" + name)
os.makedirs(obj_dir, exist_ok=True)
procs, objects = [], []
for unit, flags in _build.UNITS:
    obj = os.path.join(obj_dir, os.path.splitext(unit)[0] + ".o")
    objects.append(obj)
    cmd = ["nvcc"] + _build.NVCC_FLAGS + flags + extra + ["-I", os.path.join(ROOT, "include"), "-c",
                                                         os.path.join(ROOT, "smartpy_b200", "csrc", unit), "-o", obj]
    procs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
log = "".join(p.communicate()[0] for p in procs)
if any(p.returncode for p in procs):
    sys.exit(log)
lib = os.path.join(out_dir, "lib_{}.so".format(name))
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib] + objects)
open(os.path.join(obj_dir, "ptxas.txt"), "w").write(log)
print(lib)
