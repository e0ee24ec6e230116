"""b200_nhwc_concat (channel concat / chunk seams) against torch.cat / Tensor.chunk: pure data movement,
so every comparison is bit-exact."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("chans", [(32, 32, 32), (64, 128), (16, 16, 16, 16, 16, 16, 16, 16), (64, 80), (7, 9), (3, 8, 5)])
def test_concat_matches_torch_cat(dtype, chans):
    from improving_yolov8_cbam_swinblock_b200 import functional as Fb

    torch.manual_seed(0)
    xs = [_cl(torch.randn(3, c, 10, 12, device="cuda").to(dtype)) for c in chans]
    out = Fb.nhwc_concat(xs)
    ref = torch.cat(xs, 1)
    assert out.shape == ref.shape and out.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(out, ref)


def test_concat_of_channel_slices_and_nchw_operands():
    from improving_yolov8_cbam_swinblock_b200 import functional as Fb

    torch.manual_seed(1)
    wide = _cl(torch.randn(4, 96, 20, 20, device="cuda").bfloat16())
    a, b = wide.chunk(2, 1)                       # row-strided views (row stride 96, 48 channels each)
    c = torch.randn(4, 32, 20, 20, device="cuda").bfloat16()   # NCHW-dense operand: converted on entry
    out = Fb.nhwc_concat([b, c, a, wide[:, 8:24]])
    assert torch.equal(out, torch.cat([b, c, a, wide[:, 8:24]], 1))


def test_concat_and_chunk_gradients_match_autograd():
    from improving_yolov8_cbam_swinblock_b200 import functional as Fb

    torch.manual_seed(2)
    x = _cl(torch.randn(2, 64, 8, 8, device="cuda")).requires_grad_(True)
    w = torch.randn(2, 96, 8, 8, device="cuda")

    def run(chunk, cat):
        y0, y1 = chunk(x)
        z = cat([y0, y1, y1 * 2.0])
        (z * w).sum().backward()
        g = x.grad.clone()
        x.grad = None
        return z.detach(), g

    z_ref, g_ref = run(lambda t: t.chunk(2, 1), lambda ts: torch.cat(ts, 1))
    z, g = run(lambda t: Fb.nhwc_chunk(t, 2), Fb.nhwc_concat)
    assert torch.equal(z, z_ref)
    assert torch.equal(g, g_ref)


def test_chunk_with_unused_half_gets_zero_gradient():
    from improving_yolov8_cbam_swinblock_b200 import functional as Fb

    x = _cl(torch.randn(2, 32, 4, 4, device="cuda").bfloat16()).requires_grad_(True)
    y0, _ = Fb.nhwc_chunk(x, 2)
    y0.float().sum().backward()
    assert torch.equal(x.grad[:, :16], torch.ones_like(x.grad[:, :16]))
    assert torch.count_nonzero(x.grad[:, 16:]) == 0


def test_concat_rejects_too_many_sources_by_falling_back():
    from improving_yolov8_cbam_swinblock_b200 import functional as Fb

    xs = [_cl(torch.randn(1, 8, 4, 4, device="cuda")) for _ in range(9)]   # > 8 sources: stock torch.cat
    assert torch.equal(Fb.nhwc_concat(xs), torch.cat(xs, 1))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("shape", [(2, 3, 64, 64), (3, 1, 8, 12), (1, 4, 6, 10), (2, 2, 5, 4)])
def test_u8_to_nhwc_is_bit_identical_to_the_reference_chain(dtype, shape):
    from improving_yolov8_cbam_swinblock_b200 import functional as Fb

    torch.manual_seed(3)
    img = torch.randint(0, 256, shape, dtype=torch.uint8, device="cuda")
    out = Fb.u8_to_nhwc(img, dtype, 255.0)
    ref = (img.float() / 255).to(dtype)          # detect/train.py:100, then autocast's cast of the conv input
    assert out.is_contiguous(memory_format=torch.channels_last) or shape[1] == 1
    assert torch.equal(out, ref)
