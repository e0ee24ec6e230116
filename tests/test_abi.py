"""CPU: the C-ABI library loads, exports every symbol include/*.h declares, and reports argument errors through
return codes + b200_last_error() (argument validation runs before any CUDA call, so no GPU is needed)."""
import ctypes as C
import os
import re

from util import ROOT


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "b200_yolo_blocks.h")).read()
    return sorted(set(re.findall(r"B200_API\s+[\w\s\*]+?\b(b200_\w+)\s*\(", txt)))


def test_every_declared_symbol_is_exported_and_bound():
    from improving_yolov8_cbam_swinblock_b200 import _lib
    import improving_yolov8_cbam_swinblock_b200.functional  # noqa: F401  registers the swin entry points
    import improving_yolov8_cbam_swinblock_b200.gemm  # noqa: F401  registers the GEMM entry points

    syms = _header_symbols()
    assert len(syms) >= 20
    h = C.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(h, s), f"{s} declared in the header but not exported by libb200yolo.so"
    assert sorted(_lib.declared_symbols()) == syms, "python binding table and header disagree"
    assert _lib.lib().b200_abi_version() == 3


def test_error_convention_without_gpu():
    from improving_yolov8_cbam_swinblock_b200 import _lib

    L = _lib.lib()
    one = C.c_void_p(16)
    rc = L.b200_sppf_pool_fwd(one, one, None, 1, 8, 4, 4, 4, 0, None)  # even k
    assert rc == 1 and b"k must be odd" in L.b200_last_error()
    rc = L.b200_sppf_pool_fwd(one, one, None, 1, 7, 4, 4, 5, 1, None)  # odd C with a 16-bit dtype
    assert rc == 3
    rc = L.b200_sppf_pool_fwd(one, one, None, 1, 8, 4, 4, 5, 9, None)  # unknown dtype
    assert rc == 2
    rc = L.b200_cbam_fwd(one, one, one, one, one, None, None, None, one, 1 << 20, 1, 8, 4, 4, 1, 5, 0, 0, None)  # ksa not in {3,7}
    assert rc == 1 and b"3 or 7" in L.b200_last_error()
    rc = L.b200_swin_attn_fwd(one, one, None, 100, 49, 32, 2, 0, 0, 0, 0, 0, None)  # tokens not a multiple of L
    assert rc == 1
    rc = L.b200_swin_attn_fwd(one, one, None, 81 * 2, 81, 32, 2, 0, 0, 0, 0, 0, None)  # window too large
    assert rc == 6
    rc = L.b200_swin_attn_fwd(one, one, None, 49 * 4, 49, 32, 2, 2, 2, 7, 7, 0, None)  # shift must be < window size
    assert rc == 1 and b"shift" in L.b200_last_error()
    assert L.b200_swin_num_tokens(64, 40, 40, 7) == 64 * 36 * 49
    assert L.b200_cbam_bwd_workspace_bytes(2, 32, 4, 4, 2, 0) >= 2 * (2 * 2 * 32 + 98) * 4
    assert L.b200_cbam_stash_bytes(2, 32, 4, 4) >= 2 * (3 * 32 + 3 * 16) * 4


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from improving_yolov8_cbam_swinblock_b200 import _lib
    import pytest

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.lib()
