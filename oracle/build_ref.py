"""TEST INFRASTRUCTURE ONLY -- recipe that makes the UNMODIFIED reference runnable on the GPU box.

The reference (mazouziwissem/improving_yolov8_CBAM_SwinBlock, an Ultralytics 8.3.108 fork) is pure Python: there is
nothing to compile, "building" it means copying the package where it lies under ``/root/reference`` into the git-ignored
``oracle/_ref/`` (it then travels to the GPU box with the working tree, like the built ``.so``; no reference source enters
the repository history).  ``oracle/ref_loader.py`` imports it from ``/root/reference`` when that exists and from
``oracle/_ref`` otherwise; consumers are the checker legs only: ``bench.py --impl reference`` / ``cpu_baseline`` /
``gpu_eager_baseline`` and the tests that drive the real ``DetectionModel`` (tests/test_gpu_plugin.py).

    python oracle/build_ref.py            # copy (idempotent); prints the destination and the file count
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("B200_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
KEEP_EXT = {".py", ".yaml", ".yml", ".json", ".txt", ".cfg", ".toml"}


def build(verbose: bool = False) -> str | None:
    """Copy <SRC>/ultralytics -> oracle/_ref/ultralytics (code + yaml configs only).  Returns DST, or None if SRC is absent
    (on the GPU box: the copy made in the dev container is what is used)."""
    pkg = os.path.join(SRC, "ultralytics")
    if not os.path.isdir(pkg):
        return None
    out = os.path.join(DST, "ultralytics")
    if os.path.isdir(out):
        shutil.rmtree(out)
    n = 0
    for root, dirs, files in os.walk(pkg):
        dirs[:] = [d for d in dirs if d != "__pycache__"]
        rel = os.path.relpath(root, pkg)
        for f in files:
            if os.path.splitext(f)[1].lower() in KEEP_EXT:
                os.makedirs(os.path.join(out, rel), exist_ok=True)
                shutil.copy2(os.path.join(root, f), os.path.join(out, rel, f))
                n += 1
    with open(os.path.join(DST, "SOURCE.txt"), "w") as fh:
        fh.write(f"verbatim copy of {pkg} ({n} files; code and yaml only) made by oracle/build_ref.py -- not tracked by git\n")
    if verbose:
        print(f"{out}: {n} files")
    return DST


if __name__ == "__main__":
    r = build(verbose=True)
    if r is None:
        print(f"{SRC}/ultralytics not found: nothing copied", file=sys.stderr)
