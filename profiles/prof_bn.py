"""Fused BN+SiLU epilogue fwd+bwd at one shape inside a profiler range (for ncu).  argv: B C H W"""
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from improving_yolov8_cbam_swinblock_b200 import functional as Fb  # noqa: E402

B, C, H, W = (int(a) for a in sys.argv[1:5]) if len(sys.argv) > 4 else (64, 64, 80, 80)
bn = nn.BatchNorm2d(C, eps=1e-3, momentum=0.03).cuda().train()
x = torch.randn(B, C, H, W, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
g = torch.randn_like(x)
for _ in range(3):
    Fb.bn_act(x, bn, True).backward(g)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
Fb.bn_act(x, bn, True).backward(g)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok")
