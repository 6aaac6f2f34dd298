"""pde_opt_b200: B200-native time-stepping hot path of acoh64/pde-opt.

Mirrors the reference's solver / equation / driver API for that path only
(SURVEY.md section 8); compute runs in hand-written sm_100a CUDA kernels behind the C ABI in
include/pdeopt_b200.h."""

__version__ = "0.1.0"
