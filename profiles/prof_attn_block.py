"""The fused attention-half forward alone at the model's shape ([64,128,40,40], ws 7, bf16) for `ncu --set full`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from improving_yolov8_cbam_swinblock_b200 import functional as Fb  # noqa: E402

dev, dt, C = torch.device("cuda"), torch.bfloat16, 128
torch.manual_seed(0)
x = torch.randn(64, C, 40, 40, device=dev).to(dt).contiguous(memory_format=torch.channels_last)
g1, b1 = torch.ones(C, device=dev), torch.zeros(C, device=dev)
win, wo = torch.randn(3 * C, C, device=dev) / C ** 0.5, torch.randn(C, C, device=dev) / C ** 0.5
bi, bo = torch.randn(3 * C, device=dev), torch.randn(C, device=dev)
train = len(sys.argv) > 1 and sys.argv[1] == "train"
for _ in range(6):
    Fb.swin_attn_block_forward_raw(x, g1, b1, win, bi, wo, bo, 2, 7, train=train)
torch.cuda.synchronize()
print("ok")
