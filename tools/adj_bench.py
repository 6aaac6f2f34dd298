"""Timing of the differentiable Cahn-Hilliard rollout (fused forward keeping states + fused adjoint)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pde_opt_b200 import secondary_bench as sb
for r in sb.c2_variants(71.4, int(sys.argv[1]) if len(sys.argv) > 1 else 4096):
    print({k: v for k, v in r.items() if k != "config"}, "|", r["config"][:70])
