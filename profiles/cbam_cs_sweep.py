"""CBAM forward at the model's P5 shape for forced cluster sizes / pixel-group counts of the resident kernel
(csrc/cbam_cluster.cu: B200_CBAM_CS, B200_CBAM_GROUPS are read once per process, hence one subprocess per setting)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch

    sys.path.insert(0, os.path.dirname(HERE))
    import improving_yolov8_cbam_swinblock_b200 as P

    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    shape = (64, 256, 20, 20)
    x = torch.randn(shape, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    mod = P.CBAM()
    mod(torch.zeros(1, shape[1], 2, 2))
    mod = mod.to(dev)
    ref = None
    for train in (False, True):
        ms = []
        for _ in range(15):
            flush.zero_()
            torch.cuda._sleep(2_000_000)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if train:
                xg = x.clone().requires_grad_(True)
                s.record()
                y = mod(xg)
                e.record()
            else:
                with torch.no_grad():
                    s.record()
                    y = mod(x)
                    e.record()
            e.synchronize()
            ms.append(s.elapsed_time(e))
        ms = sorted(ms[3:])
        print(f"CS={os.environ.get('B200_CBAM_CS', 'auto'):>4s} groups={os.environ.get('B200_CBAM_GROUPS', 'auto'):>4s} "
              f"{'train(stash)' if train else 'inference':12s} {ms[len(ms) // 2] * 1e3:7.2f} us  checksum {float(y.float().sum()):.4f}", flush=True)
else:
    for cs, gr in [("", ""), ("8", "8"), ("4", ""), ("4", "8"), ("2", ""), ("2", "8"), ("16", "")]:
        env = dict(os.environ)
        if cs:
            env["B200_CBAM_CS"] = cs
        if gr:
            env["B200_CBAM_GROUPS"] = gr
        subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, check=False)
