"""Does programmatic dependent launch (csrc/common.cuh launch_k) shorten a CUDA-graph-replayed chain of small kernels?
Replays 100 fused Conv epilogues (3 kernels each: stats -> final -> apply) on a small and a mid-sized map; run once with
B200_PDL=1 and once with B200_PDL=0 (the switch is read at library load)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from improving_yolov8_cbam_swinblock_b200 import functional as Fb  # noqa: E402

N = 100
for shape in [(64, 64, 20, 20), (64, 64, 40, 40), (64, 64, 80, 80)]:
    bn = torch.nn.BatchNorm2d(shape[1]).cuda().train()
    x = torch.randn(*shape, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    s = torch.cuda.Stream()
    with torch.no_grad(), torch.cuda.stream(s):
        for _ in range(3):
            Fb.bn_act(x, bn, True)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(N):
                z = Fb.bn_act(x, bn, True)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        t_graph = e0.elapsed_time(e1) / 10 / N * 1e3
        e0.record()
        for _ in range(N):
            Fb.bn_act(x, bn, True)
        e1.record()
        torch.cuda.synchronize()
        t_eager = e0.elapsed_time(e1) / N * 1e3
    print(f"B200_PDL={os.environ.get('B200_PDL', '1')} {shape}: graph replay {t_graph:.2f} us / epilogue (3 kernels), eager {t_eager:.2f} us")
