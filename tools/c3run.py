import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pde_opt_b200 import secondary_bench as sb
print(sb.c3_gpe(int(sys.argv[1]) if len(sys.argv)>1 else 128, 16, 71.2, 1)[0])
