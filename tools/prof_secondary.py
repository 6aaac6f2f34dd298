"""One pass over the secondary kernels at bench sizes (for ncu --set full captures)."""
import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
sys.argv = ["x"]
sys.path.insert(0, "/root/repo/tools")
import bench_configs as bc
class A: envs3=128; envs4=512; n3=256; only=""; check=False; transport="auto"
which = os.environ.get("WHICH", "c3b,c4")
for w in which.split(","):
    getattr(bc, w)(A)
