import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from improving_yolov8_cbam_swinblock_b200 import functional as Fb
dev = torch.device("cuda:0"); dt = torch.bfloat16
conv = torch.nn.Conv2d(16, 16, 3, 1, 1, bias=False).to(dev)
x = torch.randn(64, 16, 160, 160, device=dev).to(dt).contiguous(memory_format=torch.channels_last)
g = torch.randn(64, 16, 160, 160, device=dev).to(dt).contiguous(memory_format=torch.channels_last)
y = Fb.conv3x3(conv, x)
for _ in range(3):
    torch.autograd.grad(y, conv.weight, g, retain_graph=True)
torch.cuda.synchronize()
