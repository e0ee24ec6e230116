"""Each hand-written kernel once (fwd + bwd) at the model's shapes, inside a profiler range (ncu --set full)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import improving_yolov8_cbam_swinblock_b200 as P  # noqa: E402
from improving_yolov8_cbam_swinblock_b200 import functional as Fb  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
dt = torch.bfloat16
torch.manual_seed(0)
cl = lambda t: t.contiguous(memory_format=torch.channels_last)  # noqa: E731
x5 = cl(torch.randn(B, 256, 20, 20, device=dev).to(dt)).requires_grad_(True)
y0 = cl(torch.randn(B, 128, 20, 20, device=dev).to(dt)).requires_grad_(True)
x4 = cl(torch.randn(B, 128, 40, 40, device=dev).to(dt)).requires_grad_(True)
cb = P.CBAM()
cb(torch.zeros(1, 256, 2, 2))
cb = cb.to(dev)
sw = P.SwinBlock(128, 2, 7).to(dev)
bn = torch.nn.BatchNorm2d(64, eps=1e-3, momentum=0.03).to(dev).train()
x3 = cl(torch.randn(B, 64, 80, 80, device=dev).to(dt)).requires_grad_(True)


def once():
    y = cb(x5)
    y.backward(torch.ones_like(y))
    for k in (5, 7):
        c = Fb.sppf_pool(y0, k)
        c.backward(torch.ones_like(c))
    with torch.autocast("cuda", dtype=dt):
        z = sw(x4)
    z.backward(torch.ones_like(z))
    e = Fb.bn_act(x3, bn, True)   # Conv epilogue at a P3 Conv shape
    e.backward(torch.ones_like(e))


for _ in range(2):
    once()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
once()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok")
