"""Kernel-only timings of the HBM-bound blocks through the public autograd wrappers, host launch latency hidden.

Each timed launch is preceded (on the same stream) by an L2 flush (256 MB write) and a ~1 ms spin kernel, so the
start event is recorded while the GPU is still busy and the Python/ctypes overhead of the call is NOT inside the
event pair.  Prints algorithmic GB/s next to a plain 1R+1W copy of the same tensor.

usage: python profiles/ktime.py [cbam|sppf|all] [--json out.json]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import improving_yolov8_cbam_swinblock_b200 as P  # noqa: E402
from improving_yolov8_cbam_swinblock_b200 import functional as Fb  # noqa: E402

dev = torch.device("cuda:0")
dt = torch.bfloat16
PEAK = 6555.2
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def ktime(fn, iters=12, do_flush=True):
    ms = []
    for _ in range(iters + 3):
        if do_flush:
            flush.zero_()
        torch.cuda._sleep(2_000_000)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        ms.append(s.elapsed_time(e))
    ms = sorted(ms[3:])
    return ms[len(ms) // 2] * 1e3  # us


def cl(t):
    return t.contiguous(memory_format=torch.channels_last)


rows = []


def rec(name, shape, us, byts):
    gbs = byts / us / 1e3
    rows.append({"kernel": name, "shape": list(shape), "us": round(us, 2), "GBps": round(gbs, 1), "frac": round(gbs / PEAK, 4)})
    print(f"{name:18s} {str(tuple(shape)):22s} {us:8.2f} us  {gbs:8.1f} GB/s  frac {gbs / PEAK:.3f}", flush=True)


def run_cbam(shapes):
    for shape in shapes:
        torch.manual_seed(0)
        x = cl(torch.randn(shape, device=dev).to(dt))
        n = x.numel()
        mod = P.CBAM()
        mod(torch.zeros(1, shape[1], 2, 2))
        mod = mod.to(dev)
        out = torch.empty_like(x)
        rec("copy(1R+1W)", shape, ktime(lambda: out.copy_(x)), 2 * n * 2)
        with torch.no_grad():
            rec("cbam_fwd", shape, ktime(lambda: mod(x)), 2 * n * 2)
        xg = x.clone().requires_grad_(True)
        y = mod(xg)
        g = torch.randn_like(y)
        rec("cbam_bwd(+fold)", shape, ktime(lambda: torch.autograd.grad(y, xg, g, retain_graph=True)), 3 * n * 2)
        del y, xg


def run_attn(ws_list=(7, 8)):
    """window attention alone (packed qkv rows in, o / gqkv out) at the model's P4 SwinBlock: [64,128,40,40], 2 heads."""
    for ws in ws_list:
        Hp = -(-40 // ws) * ws
        L, C, nh = ws * ws, 128, 2
        T = 64 * (Hp // ws) ** 2 * L
        qkv = torch.randn(T, 3 * C, device=dev).to(dt)
        go = torch.randn(T, C, device=dev).to(dt)
        o, lse = Fb.attn_forward(qkv, T, L, C, nh)
        rec(f"attn_fwd ws={ws}", (T, 3 * C), ktime(lambda: Fb.attn_forward(qkv, T, L, C, nh)), T * C * 2 * 4 + T * nh * 4)
        rec(f"attn_bwd ws={ws}", (T, 3 * C), ktime(lambda: Fb.attn_backward(qkv, o, lse, go, T, L, C, nh)),
            T * C * 2 * 7 + T * nh * 4)


def run_sppf(shapes):
    for shape in shapes:
        torch.manual_seed(0)
        y0 = cl(torch.randn(shape, device=dev).to(dt))
        n0 = y0.numel()
        for k in (5, 7):
            with torch.no_grad():
                rec(f"sppf_fwd_k{k}", shape, ktime(lambda: Fb.sppf_pool(y0, k)), 5 * n0 * 2)
            yg = y0.clone().requires_grad_(True)
            cat = Fb.sppf_pool(yg, k)
            g = torch.randn_like(cat)
            rec(f"sppf_bwd_k{k}", shape, ktime(lambda: torch.autograd.grad(cat, yg, g, retain_graph=True)), 6 * n0 * 2)
            del cat, yg


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else "all"
    B = 64
    if what in ("cbam", "all"):
        run_cbam([(B, 256, 20, 20), (B, 128, 40, 40), (B, 64, 80, 80), (B, 512, 20, 20), (B, 576, 20, 20), (B, 256, 40, 40)])
    if what in ("sppf", "all"):
        run_sppf([(B, 128, 20, 20), (B, 256, 20, 20), (B, 288, 20, 20), (B, 64, 40, 40)])
    if what in ("attn", "all"):
        run_attn()
    if "--json" in sys.argv:
        json.dump(rows, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
