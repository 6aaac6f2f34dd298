"""pde_opt_b200: B200-native time-stepping hot path of acoh64/pde-opt.

Mirrors the reference's solver / equation / driver API for that path only
(SURVEY.md section 8); compute runs in hand-written sm_100a CUDA kernels behind the C ABI in
include/pdeopt_b200.h.  Importing the package needs neither a GPU nor the built library; the
first compute call does, and fails loudly without them (there is no CPU fallback)."""

__version__ = "0.1.0"

from .domains import Domain  # noqa: E402
from .utils import check_equation_solver_compatibility, prepare_solver_params  # noqa: E402

__all__ = ["Domain", "check_equation_solver_compatibility", "prepare_solver_params", "__version__"]
