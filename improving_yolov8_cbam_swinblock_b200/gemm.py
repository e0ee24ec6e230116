"""Dense contractions of the SwinBlock (in_proj / out_proj / mlp.0 / mlp.2 and their data gradients).

16-bit activations go to the hand-written tcgen05 GEMM of ``libb200yolo.so`` (``b200_gemm_nt``, fused bias / GELU /
residual / GELU-backward epilogues) whenever the shape is one it tiles (N, K multiples of 64).  Everything else -- f32
activations, whose parity bar is rtol 1e-5, odd channel counts, and the weight-gradient contractions X^T dY -- is a
plain library GEMM through ``torch.matmul`` (cuBLAS).  Weights are f32 parameters cast to the activation dtype per
call, which is what autocast does for ``F.linear`` in the reference (trainer.py:383).
"""
from __future__ import annotations

import torch

from . import gemm_tc as tc
from ._lib import call, dtype_code, ptr, stream_ptr

USE_TCGEN05 = True  # tests flip this to compare the two GEMM paths


def _use_tc(a, n, k):
    return USE_TCGEN05 and tc.supports(a, n, k)


def _wt(w, dtype):
    return w.detach().to(dtype)


def linear(a, w, bias):
    """a[M,K] @ w[N,K]^T + bias -> [M,N]."""
    if _use_tc(a, w.shape[0], w.shape[1]):
        return tc.gemm_nt(a, w, bias, tc.EPI_BIAS)
    wt = _wt(w, a.dtype)
    return a @ wt.t() if bias is None else torch.addmm(_wt(bias, a.dtype), a, wt.t())


def linear_res(a, w, bias, res):
    """res + a @ w^T + bias."""
    if _use_tc(a, w.shape[0], w.shape[1]):
        return tc.gemm_nt(a, w, bias, tc.EPI_BIAS_RES, residual=res)
    out = torch.addmm(res, a, _wt(w, a.dtype).t())
    return out if bias is None else out.add_(_wt(bias, a.dtype))


def linear_gelu(a, w, bias):
    """(gelu(p), p) with p = a @ w^T + bias."""
    if _use_tc(a, w.shape[0], w.shape[1]):
        return tc.gemm_nt(a, w, bias, tc.EPI_BIAS_GELU, want_preact=True)
    p = linear(a, w, bias)
    h = torch.empty_like(p)
    call("b200_swin_gelu", ptr(p), None, ptr(h), p.numel(), dtype_code(p.dtype), 0, stream_ptr(p.device))
    return h, p


def matmul_nn(a, w, add=None):
    """a[M,K] @ w[K,N] (+ add): data gradients (w is the [out,in] parameter, used un-transposed)."""
    if _use_tc(a, w.shape[1], w.shape[0]):
        wt = w.detach().t().contiguous()
        if add is None:
            return tc.gemm_nt(a, wt, None, tc.EPI_BIAS)
        return tc.gemm_nt(a, wt, None, tc.EPI_BIAS_RES, residual=add)
    wt = _wt(w, a.dtype)
    return a @ wt if add is None else torch.addmm(add, a, wt)


def matmul_nn_gelu_bwd(a, w, preact):
    """(a @ w) * gelu'(preact): GELU backward fused into the mlp.2 data-gradient GEMM."""
    if _use_tc(a, w.shape[1], w.shape[0]):
        return tc.gemm_nt(a, w.detach().t().contiguous(), None, tc.EPI_MUL_GELUGRAD, residual=preact)
    g = a @ _wt(w, a.dtype)
    call("b200_swin_gelu", ptr(preact), ptr(g), ptr(g), g.numel(), dtype_code(g.dtype), 1, stream_ptr(g.device))
    return g


def matmul_tn(a, b):
    """a[T,K]^T @ b[T,N] -> (f32 [K,N], f32 [K]): weight gradient dW = dY^T X and bias gradient sum_t dY[t,:].

    tcgen05 split-K kernel with both operands MN-major (no transposes, f32 partials folded in a fixed order; the bias
    gradient is one extra MMA against ones in the same pass); library GEMM + column-sum kernel otherwise."""
    if USE_TCGEN05 and tc.splitk_supported(a, b):
        return tc.gemm_splitk(a, b, True, True, want_colsum=True)
    from .functional import _colsum

    return (a.t() @ b).to(torch.float32), _colsum(a)
