"""Kernel-only timings of the hand-written convolution kernels (first layer forward / weight gradient, narrow 3x3 weight gradient)
next to ATen's (cuDNN) for the same tensors; L2 flushed, host launch latency hidden behind a spin kernel (see ktime.py).

usage: python profiles/ktime_conv.py
"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ktime import cl, dev, dt, ktime  # noqa: E402
from improving_yolov8_cbam_swinblock_b200 import functional as Fb  # noqa: E402


def line(name, shape, us_ours, us_aten, byts):
    print(f"{name:22s} {str(shape):34s} ours {us_ours:8.1f} us ({byts / us_ours / 1e3:7.0f} GB/s)   ATen {us_aten:8.1f} us", flush=True)


B = 64
conv = torch.nn.Conv2d(3, 16, 3, 2, 1, bias=False).to(dev)
x = cl(torch.rand(B, 3, 640, 640, device=dev).to(dt))
g = cl(torch.randn(B, 16, 320, 320, device=dev).to(dt))
wl = conv.weight.detach().to(dt)
with torch.no_grad():
    line("stem fwd", (B, 3, 640, 640, 16), ktime(lambda: Fb.stem_conv(conv, x)), ktime(lambda: F.conv2d(x, wl, None, 2, 1)),
         x.numel() * 2 + g.numel() * 2)
y = Fb.stem_conv(conv, x)
bw = lambda: torch.ops.aten.convolution_backward(g, x, wl, None, [2, 2], [1, 1], [1, 1], False, [0, 0], 1, [False, True, False])  # noqa: E731
line("stem wgrad", (B, 3, 640, 640, 16), ktime(lambda: torch.autograd.grad(y, conv.weight, g, retain_graph=True)), ktime(bw),
     x.numel() * 2 + g.numel() * 2)
for (cin, cout, s, H) in [(16, 32, 2, 320), (16, 16, 1, 160)]:
    conv = torch.nn.Conv2d(cin, cout, 3, s, 1, bias=False).to(dev)
    x = cl(torch.randn(B, cin, H, H, device=dev).to(dt))
    g = cl(torch.randn(B, cout, H // s, H // s, device=dev).to(dt))
    wl = conv.weight.detach().to(dt)
    y = Fb.conv3x3(conv, x)
    bw = lambda: torch.ops.aten.convolution_backward(g, x, wl, None, [s, s], [1, 1], [1, 1], False, [0, 0], 1, [False, True, False])  # noqa: E731
    line(f"wgrad {cin}->{cout} s{s}", (B, cin, H, H), ktime(lambda: torch.autograd.grad(y, conv.weight, g, retain_graph=True)), ktime(bw),
         x.numel() * 2 + g.numel() * 2)
