"""Timeline of CTA 0 of the fused attention-half kernel (SM clock stamps written by the kernel when a debug buffer is set)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from improving_yolov8_cbam_swinblock_b200 import _lib, functional as Fb  # noqa: E402

dev, dt, Cc = torch.device("cuda"), torch.bfloat16, 128
torch.manual_seed(0)
x = torch.randn(64, Cc, 40, 40, device=dev).to(dt).contiguous(memory_format=torch.channels_last)
g1, b1 = torch.ones(Cc, device=dev), torch.zeros(Cc, device=dev)
win, wo = torch.randn(3 * Cc, Cc, device=dev) / Cc ** 0.5, torch.randn(Cc, Cc, device=dev) / Cc ** 0.5
bi, bo = torch.randn(3 * Cc, device=dev), torch.randn(Cc, device=dev)
train = len(sys.argv) > 1 and sys.argv[1] == "train"
run = lambda: Fb.swin_attn_block_forward_raw(x, g1, b1, win, bi, wo, bo, 2, 7, train=train)  # noqa: E731
for _ in range(3):
    run()
buf = torch.zeros(16 * 16, dtype=torch.int64, device=dev)
L = _lib.lib()
L.b200_debug_set_attn_block_timeline.argtypes = [C.c_void_p]
L.b200_debug_set_attn_block_timeline(C.c_void_p(buf.data_ptr()))
run()
torch.cuda.synchronize()
L.b200_debug_set_attn_block_timeline(C.c_void_p(0))
t = buf.cpu().view(16, 16)
t0 = int(t[t > 0].min())
names = ["MMA n1_ready", "MMA qkv_issued", "MMA qk_ready", "MMA p_ready", "MMA o_ready", "MMA y_issued", "LN start", "LN x_landed",
         "LN done", "FIN y_full", "FIN stored", "WRK qkv_full", "WRK qk_written", "WRK s_full", "WRK p_written", "WRK o_written"]
ev = sorted((int(t[e, i]) - t0, names[e], i) for e in range(16) for i in range(16) if int(t[e, i]) > 0)
for v, nm, i in ev:
    print(f"{v:8d} {nm:16s} {i}")
