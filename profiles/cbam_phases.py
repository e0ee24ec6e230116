"""Per-phase latency of the CBAM forward cluster kernel from %globaltimer stamps (debug hook b200_debug_cbam_prof)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import improving_yolov8_cbam_swinblock_b200 as P  # noqa: E402
from improving_yolov8_cbam_swinblock_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
names = ["start", "tma", "A", "sync1", "mlp", "B", "sync2", "tile", "conv", "gate", "sync3"]
for shape in [(64, 256, 20, 20), (64, 128, 40, 40), (64, 64, 80, 80)]:
    x = torch.randn(shape, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    mod = P.CBAM()
    mod(torch.zeros(1, shape[1], 2, 2))
    mod = mod.to(dev)
    grid = shape[0] * 16
    buf = torch.zeros(grid * 16, dtype=torch.int64, device=dev)
    L = _lib.lib()
    L.b200_debug_cbam_prof.argtypes = [C.c_void_p]
    with torch.no_grad():
        for _ in range(3):
            mod(x)
        torch.cuda.synchronize()
        L.b200_debug_cbam_prof(C.c_void_p(buf.data_ptr()))
        mod(x)
        torch.cuda.synchronize()
        L.b200_debug_cbam_prof(C.c_void_p(0))
    t = buf.view(grid, 16).cpu()
    t = t[t[:, 0] > 0][:, :len(names)].double()
    t0 = t[:, 0].min()
    print(shape, "CTAs", t.shape[0], f"kernel span {float(t[:, len(names) - 1].max() - t0) / 1e3:.1f} us; CTA start spread {float(t[:, 0].max() - t0) / 1e3:.1f} us")
    d = t[:, 1:] - t[:, :-1]
    print("   mean per-phase us:", "  ".join(f"{n}={float(d[:, i].mean()) / 1e3:.2f}" for i, n in enumerate(names[1:])))
    print("   mean CTA lifetime us:", float((t[:, len(names) - 1] - t[:, 0]).mean()) / 1e3)
