"""One CBAM forward (+ backward with argv[1]=bwd) at the model's P5 shape inside a profiler range (for ncu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import improving_yolov8_cbam_swinblock_b200 as P  # noqa: E402

bwd = len(sys.argv) > 1 and sys.argv[1] == "bwd"
shape = (64, 256, 20, 20)
dev = torch.device("cuda:0")
x = torch.randn(shape, device=dev).bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(bwd)
mod = P.CBAM()
mod(torch.zeros(1, shape[1], 2, 2))
mod = mod.to(dev)


def fn():
    if bwd:
        mod(x).backward(torch.ones_like(x))
    else:
        with torch.no_grad():
            mod(x)


for _ in range(3):
    fn()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
fn()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok")
