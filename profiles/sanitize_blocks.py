"""One forward + backward of each hot-path block at small model-shaped inputs (bf16): the workload for
`compute-sanitizer --tool memcheck|racecheck` (one tool per gpurun call, logs committed under profiles/)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import improving_yolov8_cbam_swinblock_b200 as P  # noqa: E402

dev, dt = "cuda", torch.bfloat16
torch.manual_seed(0)


def run(mod, shape, autocast=True):
    x = torch.randn(shape, device=dev).to(dt).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    with torch.autocast("cuda", dtype=dt, enabled=autocast):
        y = mod(x)
    y.backward(torch.randn_like(y))
    torch.cuda.synchronize()
    assert torch.isfinite(y).all() and torch.isfinite(x.grad).all()
    return y


cb = P.CBAM()
cb(torch.zeros(1, 256, 2, 2))
run(cb.to(dev), (2, 256, 20, 20))            # resident cluster kernel (DSMEM) + streaming backward chain
cb4 = P.CBAM()
cb4(torch.zeros(1, 128, 2, 2))
run(cb4.to(dev), (2, 128, 40, 40))           # streaming forward chain
run(P.SwinBlock(128, 2, 7).to(dev), (2, 128, 40, 40))   # tcgen05 GEMMs + attention + the fused MLP kernels (fwd + bwd)
run(P.SwinBlock(128, 2, 8).to(dev), (1, 128, 24, 16))
run(P.SwinBlock(256, 2, 7).to(dev), (1, 256, 20, 20))   # head dim 128 attention, unfused MLP path
for k in (5, 7):
    run(P.SPPF(256, 256, k).to(dev).to(memory_format=torch.channels_last), (2, 256, 20, 20))
print("sanitize workload ok")
