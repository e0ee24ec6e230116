"""torch.profiler view of one training step: which ATen ops the non-B200 GPU time belongs to."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import improving_yolov8_cbam_swinblock_b200 as P  # noqa: E402
from improving_yolov8_cbam_swinblock_b200.harness import synthetic, train  # noqa: E402

tr = train.Trainer(P.BLOCKS, "n", 80, device="cuda:0", amp_dtype=torch.bfloat16)
tr.max_boxes = 8
dev = tr.to_device(synthetic.make_batch(64, 640, 80, seed=1234))
for _ in range(3):
    tr.step(dev)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    tr.step(dev)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=45, max_name_column_width=60))
print(prof.key_averages(group_by_input_shape=True).table(sort_by="self_cuda_time_total", row_limit=40, max_name_column_width=50, max_shapes_column_width=80))

# who calls aten::copy_ / aten::cat: aggregate CUDA time by (top-level parent op, op, shape)
import collections

agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.name in ("aten::copy_", "aten::cat", "aten::add", "aten::add_", "aten::mul", "aten::sum", "aten::div", "aten::fill_", "aten::zero_"):
        p, chain = ev.cpu_parent, []
        while p is not None:
            chain.append(p.name)
            p = p.cpu_parent
        top = chain[-1] if chain else "-"
        near = chain[0] if chain else "-"
        shp = str(ev.input_shapes[:1])
        k = (ev.name, near[:40], top[:40], shp[:60])
        agg[k][0] += 1
        agg[k][1] += ev.self_device_time_total
print("---- small-op attribution (self CUDA us)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{v[1]:9.1f} us x{v[0]:3d}  {k}")
