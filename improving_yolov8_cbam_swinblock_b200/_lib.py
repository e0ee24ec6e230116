"""ctypes binding of the C-ABI in include/b200_yolo_blocks.h (the ONLY compute path of this package).

There is deliberately no fallback: if ``libb200yolo.so`` is missing or a call fails, a RuntimeError is raised.
The handle lives in this module (process-global), never on an nn.Module, so modules stay deepcopy/pickle-safe
(SURVEY.md section 8b "Ownership / lifetime").
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200yolo.so")
ABI_VERSION = 3
F32, BF16, F16 = 0, 1, 2
CBAM_FULL, CBAM_CA, CBAM_SA = 0, 1, 2

_lock = threading.Lock()
_lib = None

_vp, _i32, _sz = C.c_void_p, C.c_int32, C.c_size_t
_SIGS = {
    "b200_abi_version": (C.c_int, []),
    "b200_last_error": (C.c_char_p, []),
    "b200_launch_count": (C.c_uint64, []),
    "b200_sppf_pool_fwd": (C.c_int, [_vp, _vp, _vp] + [_i32] * 6 + [_vp]),
    "b200_sppf_pool_bwd": (C.c_int, [_vp, _vp, _vp] + [_i32] * 6 + [_vp]),
    "b200_cbam_stash_bytes": (_sz, [_i32] * 4),
    "b200_cbam_fwd_workspace_bytes": (_sz, [_i32] * 5),
    "b200_cbam_fwd": (C.c_int, [_vp] * 9 + [_sz] + [_i32] * 8 + [_vp]),
    "b200_cbam_bwd_workspace_bytes": (_sz, [_i32] * 6),
    "b200_cbam_bwd": (C.c_int, [_vp] * 13 + [_sz] + [_i32] * 8 + [_vp]),
}


def declared_symbols():
    return sorted(_SIGS)


def register(name, restype, argtypes):
    """Used by sibling modules that bind further entry points of the same library."""
    _SIGS[name] = (restype, argtypes)
    if _lib is not None:
        fn = getattr(_lib, name)
        fn.restype, fn.argtypes = restype, argtypes


def lib():
    """Load (once) and return the shared library; raises if it is absent -- there is no other compute path."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                import torch  # noqa: F401  loads libcudart.so.12 into the process first (same runtime as torch)

                if not os.path.isfile(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} not found: build it with `python -m improving_yolov8_cbam_swinblock_b200.build` "
                        "(nvcc, sm_100a). This package has no CPU or PyTorch fallback for its kernels.")
                h = C.CDLL(LIB_PATH)
                for name, (res, args) in _SIGS.items():
                    fn = getattr(h, name)  # AttributeError here = header/library mismatch
                    fn.restype, fn.argtypes = res, args
                v = h.b200_abi_version()
                if v != ABI_VERSION:
                    raise RuntimeError(f"libb200yolo ABI version {v} != expected {ABI_VERSION}")
                _lib = h
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().b200_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


_timers = None  # name -> [(start_event, end_event)] while kernel timing is enabled (bench.py roofline leg)


def enable_timing(flag: bool = True):
    """Bracket every C-ABI call with CUDA events on the launching stream (bench.py reads them back)."""
    global _timers
    _timers = {} if flag else None


def timing_summary():
    """-> {entry point: (calls, mean ms)}; synchronises."""
    import torch

    out = {}
    if _timers:
        torch.cuda.synchronize()
        for name, evs in _timers.items():
            ms = [s.elapsed_time(e) for s, e in evs]
            out[name] = (len(ms), sum(ms) / max(len(ms), 1))
    return out


def call(name: str, *args, tag: str | None = None):
    """Invoke one C-ABI entry point; raise RuntimeError(b200_last_error()) on a non-zero return code.
    ``tag`` refines the timing key (e.g. the GEMM shape) when kernel timing is enabled."""
    fn = getattr(lib(), name)
    if _timers is None:
        rc = fn(*args)
    else:
        import torch

        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        rc = fn(*args)
        e.record()
        _timers.setdefault(tag or name, []).append((s, e))
    check(rc, name)


def launch_count() -> int:
    return int(lib().b200_launch_count())


def dtype_code(t) -> int:
    import torch

    try:
        return {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16}[t]
    except KeyError:
        raise RuntimeError(f"unsupported activation dtype {t}: the B200 kernels take float32 / bfloat16 / float16")


def stream_ptr(device):
    import torch

    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())
