import sys, torch
sys.path.insert(0, '/root/repo')
import improving_yolov8_cbam_swinblock_b200 as P
from improving_yolov8_cbam_swinblock_b200 import functional as Fb
dev='cuda'; dt=torch.bfloat16
def t(fn, n=50):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    s,e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); e.synchronize()
    return s.elapsed_time(e)/n*1e3
for shape in [(64,256,20,20),(8,256,20,20),(64,128,40,40),(64,64,80,80)]:
    x = torch.randn(shape, device=dev).to(dt).contiguous(memory_format=torch.channels_last)
    cb = P.CBAM(); cb(torch.zeros(1,shape[1],2,2)); cb = cb.to(dev)
    with torch.no_grad():
        full = t(lambda: cb(x)); ca = t(lambda: cb.ca(x)); sa = t(lambda: cb.sa(x))
        cp = t(lambda: x.clone())
    print(shape, f"full {full:.1f} us  ca-only {ca:.1f}  sa-only {sa:.1f}  torch clone {cp:.1f}")
