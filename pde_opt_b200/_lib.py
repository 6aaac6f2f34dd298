"""ctypes binding of libpdeopt_b200.so (the C ABI declared in include/pdeopt_b200.h).

The product path has no CPU fallback: if the shared library is missing, loading raises
instead of silently degrading."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# PDEOPT_LIB: alternative build of the same library (kernel experiments); never a CPU fallback
LIB_PATH = os.environ.get("PDEOPT_LIB") or os.path.join(HERE, "libpdeopt_b200.so")

MAX_COEF = 16
MAX_FUSED_STEPS = 512
NCTRL = 8

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA = 0, 1, 2, 3
KIND_CH2D, KIND_AC2D, KIND_AD2D, KIND_GPE2D = 0, 1, 2, 3
DERIVS_FD, DERIVS_FOURIER = 0, 1


class PlanDesc(ctypes.Structure):
    _fields_ = [
        ("kind", ctypes.c_int32),
        ("derivs", ctypes.c_int32),
        ("nx", ctypes.c_int32),
        ("ny", ctypes.c_int32),
        ("lo_x", ctypes.c_double),
        ("lo_y", ctypes.c_double),
        ("hx", ctypes.c_double),
        ("hy", ctypes.c_double),
        ("kappa", ctypes.c_double),
        ("mu_family", ctypes.c_int32),
        ("mu_ncoef", ctypes.c_int32),
        ("mu_coef", ctypes.c_double * MAX_COEF),
        ("mob_family", ctypes.c_int32),
        ("mob_ncoef", ctypes.c_int32),
        ("mob_coef", ctypes.c_double * MAX_COEF),
    ]


class SbmDesc(ctypes.Structure):  # pdeopt_sbm_desc
    _fields_ = [("kind", ctypes.c_int32), ("nx", ctypes.c_int32), ("ny", ctypes.c_int32), ("hx", ctypes.c_double), ("hy", ctypes.c_double),
                ("kappa", ctypes.c_double)]


class GpeDesc(ctypes.Structure):
    _fields_ = [
        ("nx", ctypes.c_int32),
        ("ny", ctypes.c_int32),
        ("lo_x", ctypes.c_double),
        ("lo_y", ctypes.c_double),
        ("hx", ctypes.c_double),
        ("hy", ctypes.c_double),
        ("k", ctypes.c_double),
        ("e", ctypes.c_double),
        ("trap_factor", ctypes.c_double),
    ]


class AdDesc(ctypes.Structure):
    _fields_ = [
        ("nx", ctypes.c_int32),
        ("ny", ctypes.c_int32),
        ("lo_x", ctypes.c_double),
        ("lo_y", ctypes.c_double),
        ("hx", ctypes.c_double),
        ("hy", ctypes.c_double),
    ]


class LineGeom(ctypes.Structure):
    """pdeopt_line_geom: offset(line, idx) = (line // n_inner)*outer + (line % n_inner)*inner
    + (idx // chunk)*hi + (idx % chunk)*lo, in elements."""
    _fields_ = [
        ("n_lines", ctypes.c_int64),
        ("n_inner", ctypes.c_int64),
        ("outer", ctypes.c_int64),
        ("inner", ctypes.c_int64),
        ("chunk", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("hi", ctypes.c_int64),
        ("lo", ctypes.c_int64),
    ]


class Ch3dDesc(ctypes.Structure):
    _fields_ = [
        ("nx", ctypes.c_int32),
        ("ny", ctypes.c_int32),
        ("nz", ctypes.c_int32),
        ("hx", ctypes.c_double),
        ("hy", ctypes.c_double),
        ("hz", ctypes.c_double),
        ("kappa", ctypes.c_double),
        ("mu_family", ctypes.c_int32),
        ("mu_ncoef", ctypes.c_int32),
        ("mu_coef", ctypes.c_double * MAX_COEF),
        ("mob_family", ctypes.c_int32),
        ("mob_ncoef", ctypes.c_int32),
        ("mob_coef", ctypes.c_double * MAX_COEF),
    ]


class PdeOptError(RuntimeError):
    pass


_lib = None

EXPORTS = [
    "pdeopt_abi_version",
    "pdeopt_last_error",
    "pdeopt_plan_create",
    "pdeopt_plan_destroy",
    "pdeopt_table_len",
    "pdeopt_sifs_step_batched",
    "pdeopt_sifs_step_batched_host",
    "pdeopt_rhs_batched",
    "pdeopt_plan_set_nonfinite_flags",
    "pdeopt_nonfinite_flags",
    "pdeopt_sifs_filter_batched",
    "pdeopt_phasefield_adjoint_work_floats",
    "pdeopt_phasefield_adjoint_step",
    "pdeopt_rhs_given_mu_batched",
    "pdeopt_phasefield_adjoint_given_mu",
    "pdeopt_sbm_rhs_batched",
    "pdeopt_sifs_rollout_fwd",
    "pdeopt_sifs_rollout_bwd",
    "pdeopt_phasefield_tangent_work_floats",
    "pdeopt_phasefield_tangent_steps",
    "pdeopt_strang_step_batched",
    "pdeopt_gpe_detect_vortices",
    "pdeopt_ad_tables_len",
    "pdeopt_ad_rollout_fwd",
    "pdeopt_ad_rollout_bwd",
    "pdeopt_fft_pos_to_freq",
    "pdeopt_fft_lines",
    "pdeopt_fft_lines_imex",
    "pdeopt_fft_lines_inv_update",
    "pdeopt_fft_lines_to_peers",
    "pdeopt_push_blocks_to_peers",
    "pdeopt_push_rows_to_peers",
    "pdeopt_fft_lines_r2c",
    "pdeopt_fft_lines_c2r_update",
    "pdeopt_ch3d_rhs",
    "pdeopt_ch3d_work_floats",
    "pdeopt_ch3d_step",
    "pdeopt_ch3d_adjoint_work_floats",
    "pdeopt_ch3d_adjoint_step",
    "pdeopt_strang_lines_work_floats",
    "pdeopt_strang_lines_step_batched",
    "pdeopt_strang_lines_step_batched_light",
    "pdeopt_measure_fp32_peak",
    "pdeopt_launch_count",
]


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PdeOptError(
            f"{LIB_PATH} is missing: build it with `python -m pde_opt_b200.build` "
            "(there is no CPU fallback for the stepping path)"
        )
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, f32 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_float
    lib.pdeopt_abi_version.restype = ctypes.c_int
    lib.pdeopt_last_error.restype = ctypes.c_char_p
    lib.pdeopt_plan_create.argtypes = [ctypes.POINTER(PlanDesc), ctypes.POINTER(vp)]
    lib.pdeopt_plan_create.restype = ctypes.c_int
    lib.pdeopt_plan_destroy.argtypes = [vp]
    lib.pdeopt_plan_destroy.restype = ctypes.c_int
    lib.pdeopt_table_len.argtypes = [vp]
    lib.pdeopt_table_len.restype = ctypes.c_int64
    sig = [vp, vp, vp, i32, i32, vp, vp, vp, vp, f32, f32, vp, vp]
    lib.pdeopt_sifs_step_batched.argtypes = sig
    lib.pdeopt_sifs_step_batched.restype = ctypes.c_int
    lib.pdeopt_sifs_step_batched_host.argtypes = sig
    lib.pdeopt_sifs_step_batched_host.restype = ctypes.c_int
    lib.pdeopt_rhs_batched.argtypes = [vp, vp, vp, i32, vp, vp]
    lib.pdeopt_rhs_batched.restype = ctypes.c_int
    lib.pdeopt_sifs_filter_batched.argtypes = [vp, vp, vp, vp, i32, f32, vp, vp]
    lib.pdeopt_sifs_filter_batched.restype = ctypes.c_int
    lib.pdeopt_phasefield_adjoint_work_floats.argtypes = [vp, i32]
    lib.pdeopt_phasefield_adjoint_work_floats.restype = ctypes.c_int64
    lib.pdeopt_phasefield_adjoint_step.argtypes = [vp, vp, vp, vp, i32, f32, vp, vp, vp, vp, vp]
    lib.pdeopt_phasefield_adjoint_step.restype = ctypes.c_int
    lib.pdeopt_sifs_rollout_fwd.argtypes = [vp, vp, vp, i32, i32, vp, vp, vp, i32, vp]
    lib.pdeopt_sifs_rollout_fwd.restype = ctypes.c_int
    lib.pdeopt_sbm_rhs_batched.argtypes = [ctypes.POINTER(SbmDesc), vp, vp, vp, vp, vp, vp, vp, f32, f32, f32, vp, vp, i32, vp]
    lib.pdeopt_sbm_rhs_batched.restype = ctypes.c_int
    lib.pdeopt_push_blocks_to_peers.argtypes = [vp, vp, i32, ctypes.c_int64, ctypes.c_int64, i32, vp]
    lib.pdeopt_push_blocks_to_peers.restype = ctypes.c_int
    lib.pdeopt_push_rows_to_peers.argtypes = [vp, vp, i32, ctypes.c_int64, ctypes.c_int64, i32, ctypes.c_int64, ctypes.c_int64, i32, vp]
    lib.pdeopt_push_rows_to_peers.restype = ctypes.c_int
    lib.pdeopt_phasefield_adjoint_given_mu.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i32, f32, vp, vp, vp]
    lib.pdeopt_phasefield_adjoint_given_mu.restype = ctypes.c_int
    lib.pdeopt_rhs_given_mu_batched.argtypes = [vp, vp, vp, vp, vp, i32, vp, vp]
    lib.pdeopt_rhs_given_mu_batched.restype = ctypes.c_int
    lib.pdeopt_sifs_rollout_bwd.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp]
    lib.pdeopt_sifs_rollout_bwd.restype = ctypes.c_int
    lib.pdeopt_phasefield_tangent_work_floats.argtypes = [vp, i32, i32]
    lib.pdeopt_phasefield_tangent_work_floats.restype = ctypes.c_int64
    lib.pdeopt_phasefield_tangent_steps.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp]
    lib.pdeopt_phasefield_tangent_steps.restype = ctypes.c_int
    lib.pdeopt_strang_step_batched.argtypes = [ctypes.POINTER(GpeDesc), vp, vp, i32, i32, vp, vp, f32, f32, vp, vp]
    lib.pdeopt_strang_step_batched.restype = ctypes.c_int
    lib.pdeopt_gpe_detect_vortices.argtypes = [vp, i32, i32, i32, f32, f32, vp, vp, vp]
    lib.pdeopt_gpe_detect_vortices.restype = ctypes.c_int
    i64 = ctypes.c_int64
    lib.pdeopt_ad_tables_len.argtypes = [ctypes.POINTER(AdDesc)]
    lib.pdeopt_ad_tables_len.restype = ctypes.c_int64
    lib.pdeopt_ad_rollout_fwd.argtypes = [ctypes.POINTER(AdDesc), vp, vp, i32, i32, vp, vp, vp, i32, i32, i32, vp, i64, vp]
    lib.pdeopt_ad_rollout_fwd.restype = ctypes.c_int
    lib.pdeopt_ad_rollout_bwd.argtypes = [ctypes.POINTER(AdDesc), vp, i64, vp, vp, i32, i32, vp, vp, vp, i32, i32, i32, vp, vp]
    lib.pdeopt_ad_rollout_bwd.restype = ctypes.c_int
    gp = ctypes.POINTER(LineGeom)
    lib.pdeopt_fft_pos_to_freq.argtypes = [i32, i32]
    lib.pdeopt_fft_pos_to_freq.restype = ctypes.c_int32
    lib.pdeopt_fft_lines.argtypes = [vp, vp, i32, gp, gp, i32, i32, f32, vp]
    lib.pdeopt_fft_lines.restype = ctypes.c_int
    lib.pdeopt_fft_lines_imex.argtypes = [vp, vp, i32, gp, vp, gp, f32, f32, vp]
    lib.pdeopt_fft_lines_imex.restype = ctypes.c_int
    lib.pdeopt_fft_lines_inv_update.argtypes = [vp, i32, gp, vp, vp, gp, f32, vp]
    lib.pdeopt_fft_lines_inv_update.restype = ctypes.c_int
    lib.pdeopt_fft_lines_to_peers.argtypes = [vp, i32, gp, ctypes.POINTER(vp), i32, gp, i64, vp, gp, f32, f32, vp]
    lib.pdeopt_fft_lines_to_peers.restype = ctypes.c_int
    lib.pdeopt_fft_lines_r2c.argtypes = [vp, vp, i32, i64, vp]
    lib.pdeopt_fft_lines_r2c.restype = ctypes.c_int
    lib.pdeopt_fft_lines_c2r_update.argtypes = [vp, i32, i64, vp, vp, f32, vp]
    lib.pdeopt_fft_lines_c2r_update.restype = ctypes.c_int
    c3 = ctypes.POINTER(Ch3dDesc)
    lib.pdeopt_ch3d_rhs.argtypes = [c3, vp, vp, vp, vp, vp, i32, vp]
    lib.pdeopt_ch3d_rhs.restype = ctypes.c_int
    lib.pdeopt_ch3d_work_floats.argtypes = [c3, i32]
    lib.pdeopt_ch3d_work_floats.restype = ctypes.c_int64
    lib.pdeopt_ch3d_step.argtypes = [c3, vp, vp, i32, i32, vp, vp, vp, vp]
    lib.pdeopt_ch3d_step.restype = ctypes.c_int
    lib.pdeopt_ch3d_adjoint_work_floats.argtypes = [c3, i32]
    lib.pdeopt_ch3d_adjoint_work_floats.restype = ctypes.c_int64
    lib.pdeopt_ch3d_adjoint_step.argtypes = [c3, vp, vp, vp, i32, f32, vp, vp, vp, vp, vp]
    lib.pdeopt_ch3d_adjoint_step.restype = ctypes.c_int
    lib.pdeopt_strang_lines_work_floats.argtypes = [i32, i32, i32]
    lib.pdeopt_strang_lines_work_floats.restype = ctypes.c_int64
    lib.pdeopt_strang_lines_step_batched.argtypes = [ctypes.POINTER(GpeDesc), vp, vp, i32, i32, vp, vp, f32, f32, vp, vp, vp]
    lib.pdeopt_strang_lines_step_batched.restype = ctypes.c_int
    lib.pdeopt_plan_set_nonfinite_flags.argtypes = [vp, vp]
    lib.pdeopt_plan_set_nonfinite_flags.restype = ctypes.c_int
    lib.pdeopt_nonfinite_flags.argtypes = [vp, i32, i64, vp, vp]
    lib.pdeopt_nonfinite_flags.restype = ctypes.c_int
    lib.pdeopt_strang_lines_step_batched_light.argtypes = [ctypes.POINTER(GpeDesc), vp, vp, i32, i32, vp, vp, f32, f32, vp, vp, i64, vp, vp]
    lib.pdeopt_strang_lines_step_batched_light.restype = ctypes.c_int
    lib.pdeopt_measure_fp32_peak.argtypes = [ctypes.POINTER(ctypes.c_double), vp]
    lib.pdeopt_measure_fp32_peak.restype = ctypes.c_int
    lib.pdeopt_launch_count.restype = ctypes.c_int64
    _lib = lib
    return lib


def device_of(*tensors):
    """Context manager that makes the CUDA device of the first tensor argument current: the C ABI
    works on the current device (kernel attributes, plan scratch, stream), so every call through
    ctypes is wrapped in it."""
    import contextlib

    import torch

    for t in tensors:
        if t is not None and hasattr(t, "is_cuda") and t.is_cuda:
            return torch.cuda.device(t.device)
    return contextlib.nullcontext()


def stream_ptr(t):
    """cudaStream_t (as void*) of torch's current stream on the device of tensor `t`."""
    import torch

    dev = t.device if (t is not None and hasattr(t, "is_cuda") and t.is_cuda) else None
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def check(status):
    if status != OK:
        msg = load().pdeopt_last_error().decode()
        raise PdeOptError(f"pdeopt status {status}: {msg}")
