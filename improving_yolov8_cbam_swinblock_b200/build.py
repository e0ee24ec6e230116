"""Build libb200yolo.so in-tree for sm_100a:  ``python -m improving_yolov8_cbam_swinblock_b200.build``."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def build(verbose: bool = False, jobs: int = 8) -> str:
    r = subprocess.run(["make", "-C", os.path.join(HERE, "csrc"), f"-j{jobs}"], capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout[-4000:] + r.stderr[-8000:])
    if r.returncode:
        raise RuntimeError("building libb200yolo.so failed (see make output above)")
    return os.path.join(HERE, "libb200yolo.so")


if __name__ == "__main__":
    print(build(verbose=True))
