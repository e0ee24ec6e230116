"""ctypes binding of the hand-written tcgen05 GEMM (csrc/gemm_tc.cu, ``b200_gemm_nt``)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import call, dtype_code, ptr, stream_ptr

_lib.register("b200_gemm_nt_supported", C.c_int, [C.c_int64, C.c_int32, C.c_int32, C.c_int32])
_lib.register("b200_gemm_nt", C.c_int, [C.c_void_p] * 6 + [C.c_int64] + [C.c_int32] * 4 + [C.c_void_p])

_lib.register("b200_gemm_splitk_workspace_bytes", C.c_size_t, [C.c_int32, C.c_int32, C.c_int64])
_lib.register("b200_gemm_splitk", C.c_int, [C.c_void_p] * 5 + [C.c_size_t, C.c_int32, C.c_int32, C.c_int64] + [C.c_int32] * 3 + [C.c_void_p])

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
         stream_ptr(a.device), tag=f"b200_gemm_nt[{M}x{N}x{K},epi{epi}]")
    return (d, d2) if d2 is not None else d


def linear(a, w, bias):
    return gemm_nt(a, w, bias, EPI_BIAS)


def splitk_supported(a: torch.Tensor, b: torch.Tensor) -> bool:
    return (a.is_cuda and a.dtype in (torch.bfloat16, torch.float16) and a.dtype == b.dtype and a.is_contiguous()
            and b.is_contiguous() and all(d % 8 == 0 for d in (*a.shape, *b.shape)))


def gemm_splitk(a, b, a_mn: bool, b_mn: bool, want_colsum: bool = False):
    """f32 D[M,N] = sum_k A(m,k) B(n,k); a is [K,M] if a_mn else [M,K]; b is [K,N] if b_mn else [N,K].
    want_colsum: also return f32 [M] = sum_k A(m,k) (computed by one extra MMA against ones in the same pass)."""
    K, M = a.shape if a_mn else a.shape[::-1]
    Kb, N = b.shape if b_mn else b.shape[::-1]
    assert K == Kb, (a.shape, b.shape)
    L = _lib.lib()
    nbytes = L.b200_gemm_splitk_workspace_bytes(M, N, K)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=a.device)
    d = torch.empty((M, N), dtype=torch.float32, device=a.device)
    cs = torch.empty(M, dtype=torch.float32, device=a.device) if want_colsum else None
    call("b200_gemm_splitk", ptr(a), ptr(b), ptr(d), ptr(cs), ptr(ws), nbytes, M, N, K, int(a_mn), int(b_mn),
         dtype_code(a.dtype), stream_ptr(a.device), tag=f"b200_gemm_splitk[{M}x{N}x{K}]")
    return (d, cs) if want_colsum else d
