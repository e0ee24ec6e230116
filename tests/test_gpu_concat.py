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


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("shape,sh,sw", [((2, 64, 10, 12), 2, 2), ((3, 8, 5, 7), 2, 2), ((1, 16, 4, 6), 3, 1), ((2, 24, 3, 3), 1, 2)])
def test_nearest_upsample_forward_and_backward_match_aten(dtype, shape, sh, sw):
    import torch.nn.functional as F

    from improving_yolov8_cbam_swinblock_b200 import functional as Fb

    if (shape[1] * torch.empty(0, dtype=dtype).element_size()) % 16:
        pytest.skip("channel width below one 16-byte vector: stock op")
    torch.manual_seed(4)
    x = _cl(torch.randn(shape, device="cuda").to(dtype)).requires_grad_(True)
    out = Fb.nhwc_upsample_nearest(x, sh, sw)
    ref = F.interpolate(x.detach().float(), scale_factor=(float(sh), float(sw)), mode="nearest").to(dtype)
    assert torch.equal(out, ref)                     # pure data movement
    g = _cl(torch.randn_like(out))
    out.backward(g)
    gref = g.float().view(shape[0], shape[1], shape[2], sh, shape[3], sw).sum((3, 5))   # f32 sum of the sh*sw gradients
    tol = 0 if dtype == torch.float32 else 2.0 ** -7
    assert torch.allclose(x.grad.float(), gref.to(dtype).float(), rtol=tol, atol=1e-6 if tol == 0 else tol)


def test_nearest_upsample_backward_reads_a_concat_slice_in_place():
    from improving_yolov8_cbam_swinblock_b200 import functional as Fb

    torch.manual_seed(5)
    x = _cl(torch.randn(2, 32, 6, 6, device="cuda").bfloat16()).requires_grad_(True)
    other = _cl(torch.randn(2, 16, 12, 12, device="cuda").bfloat16())
    z = Fb.nhwc_concat([Fb.nhwc_upsample_nearest(x, 2, 2), other])
    w = _cl(torch.randn_like(z))
    (z.float() * w.float()).sum().backward()
    gref = w[:, :32].float().view(2, 32, 6, 2, 6, 2).sum((3, 5)).bfloat16()
    assert torch.allclose(x.grad.float(), gref.float(), rtol=2.0 ** -7, atol=2.0 ** -7)


# ---- gradient fan-in (b200_nhwc_add) and the fork that routes a multi-consumer map's gradients through it ---------------
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("n", [2, 3, 4])
def test_fan_in_add_of_dense_and_sliced_sources(dtype, n):
    from improving_yolov8_cbam_swinblock_b200 import functional as Fb

    torch.manual_seed(10 + n)
    wide = _cl(torch.randn(3, 160, 9, 11, device="cuda").to(dtype))
    srcs = [wide[:, 32:96], _cl(torch.randn(3, 64, 9, 11, device="cuda").to(dtype)), wide[:, 96:160], wide[:, 0:64]][:n]
    assert Fb._add_ok(srcs)
    out = Fb.nhwc_add(srcs)
    acc = srcs[0].float()
    for s in srcs[1:]:
        acc = acc + s.float()          # f32 sum in list order, one rounding
    assert out.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(out, acc.to(dtype))
    if n == 2:
        assert torch.equal(out, srcs[0] + srcs[1])   # two operands: exactly ATen's a + b


def test_fan_in_add_full_size_and_fallback():
    from improving_yolov8_cbam_swinblock_b200 import functional as Fb

    torch.manual_seed(3)
    g_cat = _cl(torch.randn(64, 96, 80, 80, device="cuda").bfloat16())     # the P3 C2f concat's gradient
    dense = _cl(torch.randn(64, 32, 80, 80, device="cuda").bfloat16())
    sl = g_cat[:, 64:96]
    assert torch.equal(Fb.nhwc_add([sl, dense]), sl + dense)
    odd = [torch.randn(2, 7, 5, 5, device="cuda"), torch.randn(2, 7, 5, 5, device="cuda")]   # 28-byte rows: stock adds
    assert not Fb._add_ok(odd)
    assert torch.equal(Fb.nhwc_add(odd), odd[0] + odd[1])


def test_fork_gradient_equals_autograd_accumulation():
    from improving_yolov8_cbam_swinblock_b200 import functional as Fb

    torch.manual_seed(4)
    x = _cl(torch.randn(4, 64, 12, 12, device="cuda").bfloat16()).requires_grad_(True)
    other = _cl(torch.randn(4, 32, 12, 12, device="cuda").bfloat16())
    w_cat = torch.randn(4, 96, 12, 12, device="cuda").bfloat16()
    w_b = _cl(torch.randn(4, 64, 12, 12, device="cuda").bfloat16())

    def run(fork):
        a, b = fork(x)
        z = Fb.nhwc_concat([other, a])          # a's gradient: a channel slice of z's gradient (row-strided view)
        ((z * w_cat).sum() + (b * w_b).sum()).backward()
        g = x.grad.clone()
        x.grad = None
        return g

    g_ref = run(lambda t: (t, t))
    g = run(lambda t: Fb.nhwc_fork(t, 2))
    assert torch.equal(g, g_ref)
    with torch.no_grad():
        a, b = Fb.nhwc_fork(x, 2)
    assert a is x and b is x                    # no gradient: plain repetition
