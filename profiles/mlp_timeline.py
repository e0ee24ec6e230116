"""Timeline of CTA 0 of the fused MLP kernel (SM clock stamps written by the kernel when a debug buffer is set)."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from improving_yolov8_cbam_swinblock_b200 import _lib, functional as Fb  # noqa: E402

dev = torch.device("cuda")
dt = torch.bfloat16
Cc, rows = 128, 64 * 1600
torch.manual_seed(0)
y1 = torch.randn(rows, Cc, device=dev).to(dt)
w1 = torch.randn(4 * Cc, Cc, device=dev) / Cc ** 0.5
w2 = torch.randn(Cc, 4 * Cc, device=dev) / (4 * Cc) ** 0.5
b1, b2 = torch.randn(4 * Cc, device=dev), torch.randn(Cc, device=dev)
gamma, beta = torch.ones(Cc, device=dev), torch.zeros(Cc, device=dev)
w1f, b1f, w2h = Fb.swin_mlp_prep(gamma, beta, w1, b1, w2, dt)
BWD = len(sys.argv) > 1 and sys.argv[1] == "bwd"
g = torch.randn(rows, Cc, device=dev).to(dt)
run = (lambda: Fb.swin_mlp_backward_raw(g, y1, w1f, b1f, w2h)) if BWD else (lambda: Fb.swin_mlp_forward_raw(y1, w1f, b1f, w2h, b2))
for _ in range(3):
    run()
buf = torch.zeros(16 * 64, dtype=torch.int64, device=dev)
L = _lib.lib()
L.b200_debug_set_mlp_timeline.argtypes = [C.c_void_p]
L.b200_debug_set_mlp_timeline(C.c_void_p(buf.data_ptr()))
run()
torch.cuda.synchronize()
L.b200_debug_set_mlp_timeline(C.c_void_p(0))
t = buf.cpu().view(4, 4, 64)
t0 = int(t[t > 0].min())
names = {0: ("MMA", ["mma1_issue", "mma1_issued", "mma2_issue", "mma2_issued"]), 1: ("LN", ["start", "y_landed", "done", "-"]),
         2: ("EPI", ["start", "o_full", "done", "-"]), 3: ("GELU", ["start", "h_full", "computed", "done"])}
ev = []
for r in range(4):
    for e in range(4):
        for i in range(64):
            v = int(t[r, e, i])
            if v > 0:
                ev.append((v - t0, names[r][0], names[r][1][e], i))
for v, role, e, i in sorted(ev)[:400]:
    print(f"{v:8d} {role:5s} {e:12s} {i}")
