"""Reward helpers of the RL environments (mirror of pde_opt/rl_utils.py).

`detect_vortices` keeps the reference's name, arguments and result keys (rl_utils.py:19-84) for one
state; batched states ([B, N, N, 2] or complex [B, N, N]) return per-environment counts.  The
winding numbers and the counts are computed by one CUDA kernel (csrc/vortex.cuh) through
`pdeopt_gpe_detect_vortices`; there is no CPU fallback.
"""
import torch

from . import _lib


def density(psi):
    """|psi|^2 (rl_utils.py:11-12); accepts complex tensors or the interleaved [..., 2] float layout."""
    psi = torch.as_tensor(psi)
    if torch.is_complex(psi):
        return psi.real**2 + psi.imag**2
    return psi[..., 0] ** 2 + psi[..., 1] ** 2


def _as_pairs(psi):
    psi = torch.as_tensor(psi)
    if torch.is_complex(psi):
        psi = torch.view_as_real(psi.to(torch.complex64))
    if psi.shape[-1] != 2:
        raise ValueError("psi must be complex or have a trailing (re, im) axis")
    if not psi.is_cuda:
        raise _lib.PdeOptError("detect_vortices runs on the GPU: pass a CUDA tensor (there is no CPU fallback)")
    return psi.to(torch.float32).contiguous()


def vortex_counts(psi, amp_thresh=0.0, tol=0.5, winding=False):
    """Batched detection: psi [B, N0, N1, 2] (or complex [B, N0, N1]) ->
    counts int32 [B, 3] = (num_vortices, total_topological_charge, abs_charge_count) and, on request,
    the winding field int32 [B, N0, N1].  Device tensors, no host synchronisation."""
    y = _as_pairs(psi)
    if y.dim() != 4:
        raise ValueError("vortex_counts expects a batch [B, N0, N1, 2]")
    B, n0, n1, _ = y.shape
    counts = torch.empty((B, 3), dtype=torch.int32, device=y.device)
    w = torch.empty((B, n0, n1), dtype=torch.int32, device=y.device) if winding else None
    _lib.check(
        _lib.load().pdeopt_gpe_detect_vortices(
            y.data_ptr(), B, n0, n1, float(amp_thresh), float(tol), w.data_ptr() if winding else None,
            counts.data_ptr(), torch.cuda.current_stream(y.device).cuda_stream,
        )
    )
    return (counts, w) if winding else counts


def detect_vortices(psi, amp_thresh=0.0, tol=0.5):
    """rl_utils.py:19-84 for one state psi (N, N) complex (or [N, N, 2]): same result keys."""
    y = _as_pairs(psi)
    if y.dim() != 3:
        raise ValueError("detect_vortices expects one state (N, N); use vortex_counts for batches")
    counts, w = vortex_counts(y[None], amp_thresh, tol, winding=True)
    w = w[0]
    idx = torch.nonzero(w)
    charges = w[w != 0]
    c = counts[0].tolist()
    return {
        "winding": w,
        "positions": idx.to(torch.float32) + 0.5,
        "charges": charges,
        "num_vortices": int(c[0]),
        "total_topological_charge": int(c[1]),
        "abs_charge_count": int(c[2]),
    }
