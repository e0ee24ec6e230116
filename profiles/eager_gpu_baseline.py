"""Comparator (NOT part of bench.py): the same blocks as plain PyTorch-eager modules on the SAME B200 -- the honest
bar for the hand-written kernels, since the reference ships no CUDA of its own.

Eager modules = oracle/modules.py (from-primitives restatement of cbam.py / swin_block.py / block.py SPPF) and, for the
SwinBlock, also torch's own nn.MultiheadAttention-based module exactly as swin_block.py:23-58 builds it.
Prints one JSON object: per-block fwd / fwd+bwd milliseconds (bf16 autocast, B=64, model shapes) for eager vs ours,
and the whole-model training step with eager blocks vs ours.
"""
import json
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import improving_yolov8_cbam_swinblock_b200 as P  # noqa: E402
from improving_yolov8_cbam_swinblock_b200.harness import graph, synthetic, train  # noqa: E402
from oracle import modules as om  # noqa: E402


class RefSwinBlock(nn.Module):
    """swin_block.py:23-58 re-typed with torch's own modules (same op sequence as the reference on a GPU)."""

    def __init__(self, dim, num_heads=2, window_size=7):
        super().__init__()
        self.window_size = window_size
        self.norm1 = nn.LayerNorm(dim)
        self.attn = nn.MultiheadAttention(dim, num_heads, batch_first=True)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = nn.Sequential(nn.Linear(dim, dim * 4), nn.GELU(), nn.Linear(dim * 4, dim))

    def forward(self, x):
        B, C, H, W = x.shape
        ws = self.window_size
        x = F.pad(x, (0, (ws - W % ws) % ws, 0, (ws - H % ws) % ws))
        Hp, Wp = x.shape[2:]
        t = x.permute(0, 2, 3, 1).reshape(B, Hp // ws, ws, Wp // ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws, C)
        t = self.norm1(t)
        a, _ = self.attn(t, t, t)
        t = t + a
        t = t + self.mlp(self.norm2(t))
        t = t.reshape(B, Hp // ws, Wp // ws, ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, C)
        return t.permute(0, 3, 1, 2)[:, :, :H, :W]


def timeit(fn, iters=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    e.synchronize()
    return s.elapsed_time(e) / iters


def block_times(mod, x, dt=torch.bfloat16):
    mod = mod.cuda()
    with torch.no_grad(), torch.autocast("cuda", dtype=dt):
        fwd = timeit(lambda: mod(x))
    xg = x.clone().requires_grad_(True)

    def fb():
        with torch.autocast("cuda", dtype=dt):
            y = mod(xg)
        y.backward(torch.ones_like(y))

    return {"fwd_ms": round(fwd, 4), "fwd_bwd_ms": round(timeit(fb), 4)}


def main():
    dev, dt, B = "cuda", torch.bfloat16, 64
    torch.manual_seed(0)
    out = {"config": "B=64, bf16 autocast, YOLOv8n-CBAM-Swin shapes, events over 20 launches (no L2 flush)"}
    x5 = torch.randn(B, 256, 20, 20, device=dev).to(dt)
    x4 = torch.randn(B, 128, 40, 40, device=dev).to(dt)
    cl = lambda t: t.contiguous(memory_format=torch.channels_last)  # noqa: E731
    def mk_cbam(cls):
        m = cls()
        m(torch.zeros(1, 256, 2, 2))
        return m
    out["cbam_P5"] = {"eager": block_times(mk_cbam(om.CBAM), x5), "ours": block_times(mk_cbam(P.CBAM), cl(x5))}
    out["swin_P4"] = {"eager_restatement": block_times(om.SwinBlock(128, 2, 7), x4),
                      "eager_torch_mha": block_times(RefSwinBlock(128, 2, 7), x4),
                      "ours": block_times(P.SwinBlock(128, 2, 7), cl(x4))}
    for k in (5, 7):
        out[f"sppf_k{k}_P5"] = {"eager": block_times(om.make_sppf(graph.Conv)(256, 256, k), x5),
                                "ours": block_times(P.SPPF(256, 256, k).to(memory_format=torch.channels_last), cl(x5))}
    # whole-model training step
    host = synthetic.make_batch(B, 640, 80, seed=1234, pin=True)
    res = {}
    for name, blocks in (("eager_blocks", {"CBAM": om.CBAM, "SwinBlock": RefSwinBlock, "SPPF": om.make_sppf(graph.Conv)}),
                         ("ours", P.BLOCKS)):
        tr = train.Trainer(blocks, "n", 80, device="cuda:0", amp_dtype=dt)
        tr.max_boxes = 8
        d = tr.to_device(host)
        ms = timeit(lambda: tr.step(d), iters=10)
        res[name] = {"ms_per_step": round(ms, 3), "img_per_s": round(B / ms * 1e3, 1)}
        del tr
        torch.cuda.empty_cache()
    out["train_step"] = res
    print(json.dumps(out))


if __name__ == "__main__":
    main()
