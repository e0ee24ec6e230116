"""-m gpu: the fused classification-loss kernels (csrc/det_loss.cu) against torch's BCEWithLogits on the dense one-hot target
(utils/loss.py:235 with the target of tal.py:98-107): value and the gradient of every class map."""
import pytest
import torch
import torch.nn.functional as F

from util import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("bs,nc,sizes", [(2, 80, [(8, 8), (4, 4), (2, 2)]), (64, 80, [(80, 80), (40, 40), (20, 20)]), (3, 16, [(5, 7)])])
def test_cls_bce_sum_matches_torch(dtype, bs, nc, sizes):
    from improving_yolov8_cbam_swinblock_b200 import functional as Fb

    torch.manual_seed(bs + nc)
    maps = [(2 * torch.randn(bs, nc, h, w, device="cuda") - 1).to(dtype).contiguous(memory_format=torch.channels_last).requires_grad_(True)
            for h, w in sizes]
    A = sum(h * w for h, w in sizes)
    label = torch.randint(-1, nc, (bs, A), device="cuda")
    label[torch.rand(bs, A, device="cuda") < 0.9] = -1
    value = torch.rand(bs, A, device="cuda") * (label >= 0)
    got = Fb.cls_bce_sum(maps, label, value)
    (got * 0.37).backward()
    x = torch.cat([m.detach().permute(0, 2, 3, 1).reshape(bs, -1, nc) for m in maps], 1).double().requires_grad_(True)
    t = F.one_hot(label.clamp(min=0), nc).double() * value.double().unsqueeze(-1)
    want = F.binary_cross_entropy_with_logits(x, t, reduction="sum")
    (want * 0.37).backward()
    assert rel_err(got, want) < 1e-5, (float(got), float(want))
    off = 0
    for m, (h, w) in zip(maps, sizes):
        g = m.grad.permute(0, 2, 3, 1).reshape(bs, h * w, nc)
        assert m.grad.is_contiguous(memory_format=torch.channels_last)
        assert rel_err(g, x.grad[:, off:off + h * w]) < (1e-6 if dtype == torch.float32 else 4e-3)
        off += h * w
    # deterministic
    assert torch.equal(got, Fb.cls_bce_sum([m.detach() for m in maps], label, value))
