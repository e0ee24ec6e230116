"""Kernel-only timing of the fused SwinBlock halves at the model's shape ([64,128,40,40], ws 7): L2 flushed between
launches, host launch latency hidden behind a spin kernel.  Prints one JSON object."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from improving_yolov8_cbam_swinblock_b200 import functional as Fb  # noqa: E402
from improving_yolov8_cbam_swinblock_b200.harness.sweep import _time  # noqa: E402


def main():
    dev = torch.device("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    dt = torch.bfloat16
    C, rows = 128, 64 * 1600
    torch.manual_seed(0)
    y1 = torch.randn(rows, C, device=dev).to(dt)
    g = torch.randn(rows, C, device=dev).to(dt)
    w1 = torch.randn(4 * C, C, device=dev) / C ** 0.5
    w2 = torch.randn(C, 4 * C, device=dev) / (4 * C) ** 0.5
    b1, b2 = torch.randn(4 * C, device=dev), torch.randn(C, device=dev)
    gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    out = {}
    w1f, b1f, w2h = Fb.swin_mlp_prep(gamma, beta, w1, b1, w2, dt)
    out["swin_mlp_prep_us"] = 1e3 * _time(lambda: Fb.swin_mlp_prep(gamma, beta, w1, b1, w2, dt), 10, flush)
    ms_t = _time(lambda: Fb.swin_mlp_forward_raw(y1, w1f, b1f, w2h, b2, save_h=True), 20, flush)
    out["swin_mlp_fwd_train_us"] = 1e3 * ms_t
    ms = _time(lambda: Fb.swin_mlp_forward_raw(y1, w1f, b1f, w2h, b2), 20, flush)
    flops = 2.0 * rows * C * 4 * C * 2
    out["swin_mlp_fwd_us"] = 1e3 * ms
    out["swin_mlp_fwd_tflops"] = flops / (ms * 1e-3) / 1e12
    if hasattr(Fb, "swin_mlp_backward_raw"):
        ms = _time(lambda: Fb.swin_mlp_backward_raw(g, y1, w1f, b1f, w2h), 20, flush)
        out["swin_mlp_bwd_us"] = 1e3 * ms
        out["swin_mlp_bwd_tflops"] = 1.5 * flops / (ms * 1e-3) / 1e12
    # attention half: x -> y1 (pixel order), one kernel
    B, H, W, ws = 64, 40, 40, 7
    x = torch.randn(B, C, H, W, device=dev).to(dt).contiguous(memory_format=torch.channels_last)
    g1, bt1 = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    win = torch.randn(3 * C, C, device=dev) / C ** 0.5
    wo = torch.randn(C, C, device=dev) / C ** 0.5
    bi, bo = torch.randn(3 * C, device=dev), torch.randn(C, device=dev)
    for train in (False, True):
        ms = _time(lambda: Fb.swin_attn_block_forward_raw(x, g1, bt1, win, bi, wo, bo, 2, ws, train=train), 20, flush)
        out["swin_attn_block_fwd_%s_us" % ("train" if train else "infer")] = 1e3 * ms
    fl = B * H * W * (8 * C * C + 4 * ws * ws * C)
    out["swin_attn_block_fwd_infer_tflops"] = fl / (out["swin_attn_block_fwd_infer_us"] * 1e-6) / 1e12
    print(json.dumps(out))


if __name__ == "__main__":
    main()
