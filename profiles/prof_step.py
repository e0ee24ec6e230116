"""One training step of the bench workload inside a cudaProfilerStart/Stop range (for `ncu --profile-from-start off`)."""
import os
import sys

import torch

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
torch.cuda.cudart().cudaProfilerStart()
tr.step(dev)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok")
