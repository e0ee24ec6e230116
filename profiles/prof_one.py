"""Run ONE hand-written kernel (by name) a few times at the model's shapes, inside a profiler range."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import improving_yolov8_cbam_swinblock_b200 as P  # noqa: E402
from improving_yolov8_cbam_swinblock_b200 import functional as Fb  # noqa: E402

which = sys.argv[1]
B = 64
dev = torch.device("cuda:0")
dt = torch.bfloat16
torch.manual_seed(0)
cl = lambda t: t.contiguous(memory_format=torch.channels_last)  # noqa: E731
if which.startswith("sppf"):
    y0 = cl(torch.randn(B, 128, 20, 20, device=dev).to(dt)).requires_grad_(True)
    fn = lambda: Fb.sppf_pool(y0, 5).backward(torch.ones(B, 512, 20, 20, device=dev, dtype=dt).contiguous(memory_format=torch.channels_last))  # noqa: E731
elif which.startswith("cbam"):
    x5 = cl(torch.randn(B, 256, 20, 20, device=dev).to(dt)).requires_grad_(True)
    cb = P.CBAM()
    cb(torch.zeros(1, 256, 2, 2))
    cb = cb.to(dev)
    fn = lambda: cb(x5).backward(torch.ones_like(x5))  # noqa: E731
else:
    x4 = cl(torch.randn(B, 128, 40, 40, device=dev).to(dt)).requires_grad_(True)
    sw = P.SwinBlock(128, 2, 7).to(dev)

    def fn():
        with torch.autocast("cuda", dtype=dt):
            z = sw(x4)
        z.backward(torch.ones_like(z))
for _ in range(2):
    fn()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
fn()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok")
