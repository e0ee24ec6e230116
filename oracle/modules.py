"""TEST INFRASTRUCTURE ONLY -- nn.Module shells around ``oracle/blocks.py`` with the reference's names.

Same constructor arguments, sub-module names and ``state_dict`` keys as the reference classes
(``cbam.py:5-71``, ``swin_block.py:23-58``, ``block.py:201-226``), forward = the from-primitives restatement.
Used by tests (as the checker for the CUDA modules), by ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs (kind="port": the reference package cannot travel to the GPU box) and nowhere else.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import blocks as ob


class ChannelAttention(nn.Module):
    def __init__(self, in_planes=None, ratio=16):
        super().__init__()
        self.in_planes, self.ratio = in_planes, ratio
        self.shared_MLP = None
        if in_planes is not None:
            self.create_mlp(in_planes)

    def create_mlp(self, in_planes):  # cbam.py:20-27
        r = max(1, in_planes // self.ratio)
        self.shared_MLP = nn.Sequential(nn.Conv2d(in_planes, r, 1, bias=False), nn.ReLU(),
                                        nn.Conv2d(r, in_planes, 1, bias=False))

    def weights(self, x):
        if self.shared_MLP is None:  # lazy creation at first forward, cbam.py:31-33
            self.create_mlp(x.shape[1])
        w1, w2 = self.shared_MLP[0].weight, self.shared_MLP[2].weight
        return w1.reshape(w1.shape[0], -1), w2.reshape(w2.shape[0], -1)

    def forward(self, x):
        w1, w2 = self.weights(x)
        return ob.cbam_channel_attention(x, w1, w2)[:, :, None, None]


class SpatialAttention(nn.Module):
    def __init__(self, kernel_size=7):
        super().__init__()
        assert kernel_size in (3, 7), "kernel size must be 3 or 7"
        self.conv = nn.Conv2d(2, 1, kernel_size, padding=3 if kernel_size == 7 else 1, bias=False)

    def forward(self, x):
        return ob.cbam_spatial_attention(x, self.conv.weight)[:, None]


class CBAM(nn.Module):
    def __init__(self, channels=None):
        super().__init__()
        self.ca = ChannelAttention(channels, ratio=8 if channels and channels < 128 else 16)  # cbam.py:59
        self.sa = SpatialAttention(kernel_size=7)

    def forward(self, x):
        w1, w2 = self.ca.weights(x)
        return ob.cbam_forward(x, w1, w2, self.sa.conv.weight)


class _InProj(nn.Module):
    """Parameter holder exposing nn.MultiheadAttention's key names (in_proj_weight, in_proj_bias, out_proj.*)."""

    def __init__(self, dim):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * dim, dim))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * dim))
        self.out_proj = nn.Linear(dim, dim)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.constant_(self.out_proj.bias, 0.0)


class SwinBlock(nn.Module):
    def __init__(self, dim, num_heads=2, window_size=7, shift_size=0):
        super().__init__()
        self.dim, self.num_heads, self.window_size, self.shift_size = dim, num_heads, window_size, shift_size
        self.norm1 = nn.LayerNorm(dim)
        self.attn = _InProj(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = nn.Sequential(nn.Linear(dim, dim * 4), nn.GELU(), nn.Linear(dim * 4, dim))

    def forward(self, x):
        p = {k: v for k, v in self.named_parameters()}
        return ob.swin_forward(x, p, self.num_heads, self.window_size, self.shift_size)


def make_sppf(conv_cls):
    """SPPF bound to the harness' stock ``Conv`` class (cv1/cv2 stay stock)."""

    class SPPF(nn.Module):
        def __init__(self, c1, c2, k=5):
            super().__init__()
            c_ = c1 // 2
            self.cv1 = conv_cls(c1, c_, 1, 1)
            self.cv2 = conv_cls(c_ * 4, c2, 1, 1)
            self.k = k

        def forward(self, x):
            return self.cv2(ob.sppf_pool_cascade(self.cv1(x), self.k)[0])

    return SPPF
