"""Binding of the tcgen05 GEMM entry points (filled in once csrc/gemm_tc.cu is validated on hardware)."""
from __future__ import annotations


def supports(a, n, k) -> bool:
    return False


def linear(a, w, bias):  # pragma: no cover
    raise RuntimeError("tcgen05 GEMM not built")
