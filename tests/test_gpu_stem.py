"""-m gpu: the first layer's convolution Conv2d(3, c2, 3, 2, 1, bias=False) (csrc/stem_conv.cu) against F.conv2d in fp32 on the same
(rounded) image and weight: output and weight gradient; module hook falls back to the stock convolution for anything else."""
import pytest
import torch
import torch.nn.functional as F

from util import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype,tol", [(torch.bfloat16, 1e-2), (torch.float16, 2e-3)])
@pytest.mark.parametrize("B,H,W,c2", [(2, 64, 64, 16), (3, 70, 96, 16), (2, 128, 160, 32), (1, 32, 32, 48), (8, 640, 640, 16)])
def test_stem_conv_matches_conv2d(dtype, tol, B, H, W, c2):
    from improving_yolov8_cbam_swinblock_b200 import _lib, functional as Fb

    torch.manual_seed(B + H + c2)
    conv = torch.nn.Conv2d(3, c2, 3, 2, 1, bias=False).cuda()
    x = torch.rand(B, 3, H, W, device="cuda").to(dtype).contiguous(memory_format=torch.channels_last)
    g = torch.randn(B, c2, H // 2, W // 2, device="cuda").to(dtype).contiguous(memory_format=torch.channels_last)
    n0 = _lib.launch_count()
    y = Fb.stem_conv(conv, x)
    y.backward(g)
    assert _lib.launch_count() - n0 >= 2, "stem_conv did not run the hand-written kernels"
    assert y.is_contiguous(memory_format=torch.channels_last) and y.dtype == dtype
    gw = conv.weight.grad.clone()
    w32 = conv.weight.detach().to(dtype).float().requires_grad_(True)   # the kernel consumes the rounded weight (as autocast does)
    y32 = F.conv2d(x.float(), w32, None, 2, 1)
    y32.backward(g.float())
    assert rel_err(y.float(), y32) <= tol, rel_err(y.float(), y32)
    assert rel_err(gw, w32.grad) <= 1e-4, rel_err(gw, w32.grad)          # bf16 products are exact in f32; only the order differs
    # deterministic
    conv.weight.grad = None
    y2 = Fb.stem_conv(conv, x)
    y2.backward(g)
    assert torch.equal(y2, y) and torch.equal(conv.weight.grad, gw)


def test_stem_conv_hook_falls_back():
    from improving_yolov8_cbam_swinblock_b200 import _lib, functional as Fb

    x = torch.rand(2, 3, 64, 64, device="cuda")
    for conv, inp in [(torch.nn.Conv2d(3, 16, 3, 2, 1, bias=False).cuda(), x),                      # f32, no autocast
                      (torch.nn.Conv2d(3, 24, 3, 2, 1, bias=False).cuda(), x.bfloat16()),          # width the kernel does not tile
                      (torch.nn.Conv2d(3, 16, 3, 1, 1, bias=False).cuda(), x.bfloat16()),          # stride 1
                      (torch.nn.Conv2d(3, 16, 3, 2, 1, bias=True).cuda(), x.bfloat16())]:          # bias
        n0 = _lib.launch_count()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=inp.dtype != torch.float32):
            y = Fb.stem_conv(conv, inp)
        assert _lib.launch_count() == n0
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=inp.dtype != torch.float32):
            assert torch.equal(y, conv(inp))
    # under autocast an f32 image is cast first, exactly as F.conv2d would
    conv = torch.nn.Conv2d(3, 16, 3, 2, 1, bias=False).cuda()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = Fb.stem_conv(conv, x)
        want = conv(x)
    assert y.dtype == torch.bfloat16 and rel_err(y.float(), want.float()) < 1e-2


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B,H,W,cin,cout,stride", [(2, 32, 32, 16, 16, 1), (3, 38, 64, 16, 32, 2), (2, 20, 48, 16, 32, 1), (2, 64, 96, 16, 16, 2),
                                                   (1, 6, 16, 16, 16, 1), (8, 160, 160, 16, 16, 1), (8, 320, 320, 16, 32, 2)])
def test_conv3x3_wgrad_matches_conv2d(dtype, B, H, W, cin, cout, stride):
    """csrc/conv_wgrad.cu against autograd of F.conv2d in fp32 on the same rounded tensors; output and input gradient are ATen's."""
    from improving_yolov8_cbam_swinblock_b200 import _lib, functional as Fb

    torch.manual_seed(H + cin + cout)
    conv = torch.nn.Conv2d(cin, cout, 3, stride, 1, bias=False).cuda()
    x = torch.randn(B, cin, H, W, device="cuda").to(dtype).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    g = torch.randn(B, cout, H // stride, W // stride, device="cuda").to(dtype).contiguous(memory_format=torch.channels_last)
    n0 = _lib.launch_count()
    y = Fb.conv3x3(conv, x)
    y.backward(g)
    assert _lib.launch_count() - n0 >= 1, "conv3x3 did not run the hand-written weight-gradient kernel"
    gw, gx = conv.weight.grad.clone(), x.grad.clone()
    w32 = conv.weight.detach().to(dtype).float().requires_grad_(True)
    x32 = x.detach().float().requires_grad_(True)
    y32 = F.conv2d(x32, w32, None, stride, 1)
    y32.backward(g.float())
    assert rel_err(gw, w32.grad) <= 1e-4, rel_err(gw, w32.grad)
    assert rel_err(y.float(), y32) <= (1e-2 if dtype == torch.bfloat16 else 2e-3)
    assert rel_err(gx.float(), x32.grad) <= (1e-2 if dtype == torch.bfloat16 else 2e-3)
    conv.weight.grad = None
    Fb.conv3x3(conv, x.detach()).backward(g)
    assert torch.equal(conv.weight.grad, gw)   # deterministic
    # layers the kernel does not serve keep the stock path (other widths: cuDNN's sm100 wgrad kernels are the faster ones there)
    for other, inp in [(torch.nn.Conv2d(cin, 24, 3, stride, 1, bias=False).cuda(), x.detach()),
                       (torch.nn.Conv2d(32, 32, 3, stride, 1, bias=False).cuda(), torch.cat((x.detach(), x.detach()), 1))]:
        n0 = _lib.launch_count()
        with torch.autocast("cuda", dtype=dtype):
            Fb.conv3x3(other, inp).float().sum().backward()
        assert _lib.launch_count() == n0 and other.weight.grad is not None


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B,H,W", [(2, 32, 32), (3, 38, 64), (1, 2, 32), (2, 66, 160), (16, 320, 320)])
@pytest.mark.parametrize("cl_weight", [False, True])
def test_conv3x3_dgrad_s2_matches_aten(dtype, B, H, W, cl_weight):
    """csrc/conv_dgrad.cu (input gradient of the 16 -> 32 stride-2 layer) against ATen's convolution_backward on the same 16-bit
    tensors and against the fp32 gradient: no further from fp32 than cuDNN is, ragged row tiles (H/2 % 4 != 0), a one-row map,
    both weight layouts (the harness keeps parameters channels_last), deterministic."""
    from improving_yolov8_cbam_swinblock_b200 import _lib, functional as Fb

    torch.manual_seed(H + W)
    conv = torch.nn.Conv2d(16, 32, 3, 2, 1, bias=False).cuda()
    if cl_weight:
        conv = conv.to(memory_format=torch.channels_last)
    x = torch.randn(B, 16, H, W, device="cuda").to(dtype).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    g = torch.randn(B, 32, H // 2, W // 2, device="cuda").to(dtype).contiguous(memory_format=torch.channels_last)

    def grad(mine):
        Fb.DGRAD_S2[0] = mine
        try:
            n0 = _lib.launch_count()
            Fb.conv3x3(conv, x).backward(g)
            n = _lib.launch_count() - n0
        finally:
            Fb.DGRAD_S2[0] = True
        gx = x.grad.clone()
        x.grad = None
        return gx, n

    gx, n_mine = grad(True)
    gx_aten, n_aten = grad(False)
    assert n_mine == n_aten + 1, "the hand-written input-gradient kernel did not run"
    assert gx.is_contiguous(memory_format=torch.channels_last) and gx.dtype == dtype
    w32 = conv.weight.detach().to(dtype).float()
    ref = torch.nn.grad.conv2d_input(x.shape, w32, g.float(), stride=2, padding=1)
    e_mine, e_aten = rel_err(gx.float(), ref), rel_err(gx_aten.float(), ref)
    assert e_mine <= max(1.05 * e_aten, 1e-6), (e_mine, e_aten)
    assert rel_err(gx.float(), gx_aten.float()) <= (8e-3 if dtype == torch.bfloat16 else 1e-3)
    assert torch.equal(grad(True)[0], gx)   # deterministic


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B,H,W", [(2, 32, 32), (3, 38, 64), (1, 2, 32), (2, 66, 160), (16, 320, 320)])
@pytest.mark.parametrize("cl_weight", [False, True])
def test_conv3x3_fwd_s2_matches_aten(dtype, B, H, W, cl_weight):
    """csrc/conv_dgrad.cu conv3_fwd_s2_kernel (forward of the 16 -> 32 stride-2 layer) against F.conv2d on the same 16-bit tensors and
    against the fp32 convolution: no further from fp32 than cuDNN is; ragged row tiles (H/2 odd), a one-row map, both weight layouts."""
    from improving_yolov8_cbam_swinblock_b200 import _lib, functional as Fb

    torch.manual_seed(H * 3 + W)
    conv = torch.nn.Conv2d(16, 32, 3, 2, 1, bias=False).cuda()
    if cl_weight:
        conv = conv.to(memory_format=torch.channels_last)
    x = torch.randn(B, 16, H, W, device="cuda").to(dtype).contiguous(memory_format=torch.channels_last).requires_grad_(True)

    def fwd(mine):
        Fb.FWD_S2[0] = mine
        try:
            n0 = _lib.launch_count()
            y = Fb.conv3x3(conv, x)
            return y.detach(), _lib.launch_count() - n0
        finally:
            Fb.FWD_S2[0] = True

    y, n_mine = fwd(True)
    y_aten, n_aten = fwd(False)
    assert n_mine == n_aten + 1, "the hand-written forward kernel did not run"
    assert y.shape == (B, 32, H // 2, W // 2) and y.dtype == dtype and y.is_contiguous(memory_format=torch.channels_last)
    ref = F.conv2d(x.detach().float(), conv.weight.detach().to(dtype).float(), None, 2, 1)
    e_mine, e_aten = rel_err(y.float(), ref), rel_err(y_aten.float(), ref)
    assert e_mine <= max(1.05 * e_aten, 1e-6), (e_mine, e_aten)
    assert rel_err(y.float(), y_aten.float()) <= (8e-3 if dtype == torch.bfloat16 else 1e-3)
    assert torch.equal(fwd(True)[0], y)   # deterministic
