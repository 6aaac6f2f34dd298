#!/usr/bin/env python
"""Build a variant of libpdeopt_b200.so for kernel experiments: recompiles the named csrc/*.cu files with extra
nvcc flags and links them with the default objects into tools/_exp/lib_<name>.so (used through PDEOPT_LIB).
usage: tools/build_variant.py <name> <file.cu>[,<file.cu>...] [nvcc flags...]"""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pde_opt_b200 import build as B

name, files, flags = sys.argv[1], sys.argv[2].split(","), sys.argv[3:]
B.build()
out_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_exp")
os.makedirs(out_dir, exist_ok=True)
objs = []
for src in B._sources():
    base = os.path.basename(src)
    if base in files:
        o = os.path.join(out_dir, f"{name}_{base[:-3]}.o")
        subprocess.check_call(["nvcc", *B.NVCC_FLAGS, *flags, "-c", src, "-o", o], cwd=B.CSRC)
        objs.append(o)
    else:
        objs.append(B._obj(src))
lib = os.path.join(out_dir, f"lib_{name}.so")
subprocess.check_call(["nvcc", "--shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a", *objs, "-o", lib, "-lcudart"])
print(lib)
