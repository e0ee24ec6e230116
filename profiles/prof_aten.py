"""Which stock ATen kernels are left in one eagerly launched training step, by operator and input shapes (torch.profiler).

usage: python profiles/prof_aten.py [batch] > profiles/aten_ops_rNN.txt
"""
import collections
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import improving_yolov8_cbam_swinblock_b200 as P  # noqa: E402
from improving_yolov8_cbam_swinblock_b200.harness import synthetic, train  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
tr = train.Trainer(P.BLOCKS, "n", 80, device="cuda:0", amp_dtype=torch.bfloat16)
tr.max_boxes = 8
dev = tr.to_device(synthetic.make_batch(B, 640, 80, seed=1234))
for _ in range(3):
    tr.step(dev)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    tr.step(dev)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0.0, 0])
total = 0.0
for ev in prof.events():
    t = ev.self_device_time_total
    if t <= 0 or ev.device_type != torch.autograd.DeviceType.CPU:
        continue
    key = (ev.name, str(ev.input_shapes)[:150])
    agg[key][0] += t
    agg[key][1] += 1
    total += t
print(f"total self device time {total / 1e3:.2f} ms")
byop = collections.defaultdict(lambda: [0.0, 0])
for (name, _), (t, n) in agg.items():
    byop[name][0] += t
    byop[name][1] += n
print("--- by operator")
for name, (t, n) in sorted(byop.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"{t:9.1f} us {n:5d}  {name}")
print("--- by operator and shapes")
for (name, shp), (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:90]:
    print(f"{t:9.1f} us {n:4d}  {name:42s} {shp}")
