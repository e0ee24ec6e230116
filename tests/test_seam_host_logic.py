"""Host-side logic of the seam wrappers (functional.py: nhwc_concat / nhwc_chunk / nhwc_upsample_nearest), CPU only: the stride
classifier that decides whether a tensor can be handed to b200_nhwc_concat in place, and the stock-op routes taken by tensors the
kernels do not tile (no compute calls into the library here)."""
import torch

from improving_yolov8_cbam_swinblock_b200 import functional as Fb


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last)


def test_row_stride_classifier():
    dense = _cl(torch.zeros(2, 16, 5, 7))
    assert Fb._row_strided(dense) == 16
    a, b = dense.chunk(2, 1)                                   # channel slices of a channels_last map: row stride = parent C
    assert Fb._row_strided(a) == 16 and Fb._row_strided(b) == 16
    assert Fb._row_strided(dense[:, 3:11]) == 16
    assert Fb._row_strided(torch.zeros(2, 16, 5, 7)) is None   # NCHW-dense
    assert Fb._row_strided(dense[:, :, ::2]) is None           # spatially strided view
    assert Fb._row_strided(dense[:, ::2]) is None              # channel stride 2
    assert Fb._row_strided(torch.ones(1).expand(2, 16, 5, 7)) is None   # broadcast gradient (stride 0)
    one = _cl(torch.zeros(1, 8, 1, 1))
    assert Fb._row_strided(one) == 8


def test_cpu_tensors_take_the_stock_ops():
    torch.manual_seed(0)
    x = torch.randn(2, 8, 4, 6, requires_grad=True)
    y0, y1 = Fb.nhwc_chunk(x, 2)
    z = Fb.nhwc_concat([y0, y1, y1 * 2])
    up = Fb.nhwc_upsample_nearest(z, 2, 2)
    ref = torch.nn.functional.interpolate(torch.cat([x[:, :4], x[:, 4:], x[:, 4:] * 2], 1), scale_factor=2.0, mode="nearest")
    assert torch.equal(up, ref)
    up.sum().backward()
    assert x.grad is not None and torch.isfinite(x.grad).all()


def test_concat_eligibility():
    a = torch.zeros(1, 8, 2, 2)
    assert not Fb._concat_ok([a])                              # CPU tensor
    assert not Fb._concat_ok([a.to(torch.int32)])
