"""ctypes binding of the hand-written tcgen05 GEMM (csrc/gemm_tc.cu, ``b200_gemm_nt``)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import call, dtype_code, ptr, stream_ptr

_lib.register("b200_gemm_nt_supported", C.c_int, [C.c_int64, C.c_int32, C.c_int32, C.c_int32])
_lib.register("b200_gemm_nt", C.c_int, [C.c_void_p] * 6 + [C.c_int64] + [C.c_int32] * 4 + [C.c_void_p])

EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RES, EPI_MUL_GELUGRAD = 0, 1, 2, 3


def supports(a: torch.Tensor, n: int, k: int) -> bool:
    return (a.is_cuda and a.dtype in (torch.bfloat16, torch.float16) and a.dim() == 2 and a.is_contiguous()
            and n % 64 == 0 and k % 64 == 0 and a.shape[0] >= 1 and a.data_ptr() % 16 == 0)


def gemm_nt(a, w, bias=None, epi=EPI_BIAS, residual=None, want_preact=False):
    """a[M,K] @ w[N,K]^T (+bias) with the fused epilogue; returns D or (D, preact)."""
    M, K = a.shape
    N = w.shape[0]
    wt = w.detach().to(a.dtype).contiguous()
    bf = None if bias is None else bias.detach().to(torch.float32).contiguous()
    d = torch.empty((M, N), dtype=a.dtype, device=a.device)
    d2 = torch.empty_like(d) if (epi == EPI_BIAS_GELU and want_preact) else None
    if residual is not None:
        residual = residual.contiguous()
    call("b200_gemm_nt", ptr(a), ptr(wt), ptr(bf), ptr(d), ptr(d2), ptr(residual), M, N, K, dtype_code(a.dtype), epi,
         stream_ptr(a.device))
    return (d, d2) if d2 is not None else d


def linear(a, w, bias):
    return gemm_nt(a, w, bias, EPI_BIAS)
