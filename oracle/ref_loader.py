"""TEST INFRASTRUCTURE ONLY -- loader for the real reference modules (when /root/reference exists).

The reference (mazouziwissem/improving_yolov8_CBAM_SwinBlock, an Ultralytics 8.3.108 fork) is pure
Python, so it can be imported in the dev container to (a) validate the restatement in
``oracle/blocks.py`` and (b) generate the committed fixtures under ``tests/golden``
(``oracle/make_golden.py``).  ``/root/reference`` does NOT exist on the GPU box; what travels there is the
verbatim, git-ignored copy ``oracle/_ref`` made by ``oracle/build_ref.py`` (the "compiled reference" of this
Python project).  Only checker code uses it: tests, ``bench.py --impl reference`` and its baseline legs.

Recipe follows SURVEY.md section 8(c): ``cbam.py`` / ``swin_block.py`` import only torch + einops, so they
are loaded by file path; the full ``ultralytics`` package needs permissive ``matplotlib`` stubs.
"""
import importlib.util
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root():
    """/root/reference in the dev container; on the GPU box the verbatim copy oracle/build_ref.py made (oracle/_ref)."""
    for cand in (os.environ.get("B200_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "ultralytics/nn/modules/cbam.py")):
            return cand
    return "/root/reference"


REF_ROOT = _find_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "ultralytics/nn/modules/cbam.py"))


def _load_by_path(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_cbam():
    """ultralytics/nn/modules/cbam.py:5-71 (ChannelAttention, SpatialAttention, CBAM)."""
    return _load_by_path("_ref_cbam", "ultralytics/nn/modules/cbam.py")


def load_swin():
    """ultralytics/nn/modules/swin_block.py:8-58 (window_partition, window_reverse, SwinBlock)."""
    return _load_by_path("_ref_swin_block", "ultralytics/nn/modules/swin_block.py")


class _Stub(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Stub(self.__name__ + "." + k)

    def __call__(self, *a, **k):
        return _Stub(self.__name__ + "()")


def import_ultralytics():
    """Import the full reference package (needed for SPPF/Conv/DetectionModel/loss)."""
    import torch  # noqa: F401  (import torch first so ultralytics does not pin OMP_NUM_THREADS=1)

    for m in ["matplotlib", "matplotlib.pyplot", "matplotlib.figure", "matplotlib.backends",
              "matplotlib.backends.backend_agg", "matplotlib.image", "matplotlib.font_manager",
              "matplotlib.colors", "matplotlib.patches", "matplotlib.ticker"]:
        sys.modules.setdefault(m, _Stub(m))
    os.environ.setdefault("YOLO_CONFIG_DIR", "/tmp/b200_yolo_cfg")
    os.makedirs(os.environ["YOLO_CONFIG_DIR"], exist_ok=True)
    os.environ.setdefault("YOLO_OFFLINE", "1")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import ultralytics  # noqa: F401
    return ultralytics
