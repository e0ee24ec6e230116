"""-m gpu: CBAM / ChannelAttention / SpatialAttention modules (C-ABI cluster kernels) vs the reference-generated
fixtures and the oracle.  Bars: fp32 rtol 1e-5; bf16/f16 <= 2e-2 relative error vs the fp32 oracle, outputs AND
gradients (BASELINE.json north_star)."""
import pytest
import torch

from util import assert_close_f32, load_golden, rel_err, state_dict_of, to_cl

pytestmark = pytest.mark.gpu


def _run_module(mod, x, gy):
    x = x.clone().requires_grad_(True)
    y = mod(x)
    y.backward(gy)
    return y.detach(), x.grad.detach(), {k: p.grad.detach() for k, p in mod.named_parameters() if p.grad is not None}


@pytest.mark.parametrize("name,ctor", [
    ("cbam_lazy_c32", lambda M: M.CBAM()),
    ("cbam_c64_r8", lambda M: M.CBAM(64)),
    ("cbam_lazy_c256_p5", lambda M: M.CBAM()),
    ("cbam_ca_c32", lambda M: M.ChannelAttention(32, 16)),
    ("cbam_sa_k3", lambda M: M.SpatialAttention(3)),
])
def test_golden_fp32(name, ctor):
    import improving_yolov8_cbam_swinblock_b200.modules as M

    g = load_golden(name)
    mod = ctor(M)
    x = torch.from_numpy(g["x"])
    mod(torch.zeros_like(x))  # CPU shape probe creates the lazy MLP exactly like the reference's stride pass
    mod.load_state_dict(state_dict_of(g))
    mod = mod.cuda()
    y, gx, gw = _run_module(mod, to_cl(x.cuda()), torch.from_numpy(g["gy"]).cuda())
    assert_close_f32(y, torch.from_numpy(g["y"]), name + " y")
    assert_close_f32(gx, torch.from_numpy(g["gx"]), name + " gx")
    for k, v in gw.items():
        assert_close_f32(v, torch.from_numpy(g["gw." + k]), name + " gw " + k)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.bfloat16, 2e-2), (torch.float16, 4e-3)])
# (4,256,20,20), (2,64,40,40) and the odd-sized (3,128,13,11), (2,256,9,8), (5,64,17,19) fit one SM's shared memory: in 16-bit
# dtypes they run the one-CTA-per-image kernel (csrc/cbam_image.cu; 32 / 8 / 16 lanes per pixel), the others the cluster kernel or
# the streaming chain
@pytest.mark.parametrize("shape", [(4, 256, 20, 20), (2, 64, 40, 40), (2, 128, 80, 80), (3, 48, 9, 7), (1, 576, 20, 20),
                                   (3, 128, 13, 11), (2, 256, 9, 8), (5, 64, 17, 19)])
def test_vs_oracle(dtype, tol, shape):
    """Same seeded input through the CUDA module and the fp64 oracle; 16-bit error is measured vs the oracle."""
    import improving_yolov8_cbam_swinblock_b200.modules as M
    from oracle import blocks as ob

    torch.manual_seed(sum(shape))
    B, C, H, W = shape
    mod = M.CBAM()
    mod(torch.zeros(1, C, 2, 2))
    w1 = mod.ca.shared_MLP[0].weight.detach().double().view(-1, C).requires_grad_(True)
    w2 = mod.ca.shared_MLP[2].weight.detach().double().view(C, -1).requires_grad_(True)
    ws = mod.sa.conv.weight.detach().double().requires_grad_(True)
    x = torch.randn(shape).to(dtype)
    gy = torch.randn(shape).to(dtype)
    xo = x.double().requires_grad_(True)
    yo = ob.cbam_forward(xo, w1, w2, ws)
    yo.backward(gy.double())
    mod = mod.cuda()
    xc = to_cl(x.cuda()).requires_grad_(True)
    y = mod(xc)
    assert y.dtype == dtype and y.shape == x.shape
    y.backward(to_cl(gy.cuda()))
    assert rel_err(y.cpu(), yo) <= tol
    assert rel_err(xc.grad.cpu(), xo.grad) <= tol
    assert rel_err(mod.ca.shared_MLP[0].weight.grad.cpu().view(-1, C), w1.grad) <= max(tol, 1e-5)
    assert rel_err(mod.ca.shared_MLP[2].weight.grad.cpu().view(C, -1), w2.grad) <= max(tol, 1e-5)
    assert rel_err(mod.sa.conv.weight.grad.cpu(), ws.grad) <= max(tol, 1e-5)


def test_nchw_input_and_determinism():
    import improving_yolov8_cbam_swinblock_b200.modules as M

    torch.manual_seed(3)
    mod = M.CBAM(64).cuda()
    x = torch.randn(2, 64, 12, 12, device="cuda")  # NCHW-contiguous: converted on entry
    a, b = mod(x), mod(to_cl(x))
    assert torch.equal(a, b) and a.is_contiguous(memory_format=torch.channels_last)
    g = torch.randn_like(x)
    outs = []
    for _ in range(2):
        xx = x.clone().requires_grad_(True)
        mod.zero_grad()
        mod(xx).backward(g)
        outs.append((xx.grad.clone(), mod.sa.conv.weight.grad.clone(), mod.ca.shared_MLP[0].weight.grad.clone()))
    for u, v in zip(*outs):
        assert torch.equal(u, v)  # fixed-order reductions: run-to-run reproducible


def test_half_module_and_autocast():
    """trainer/validator do .half() on the module and run under autocast."""
    import improving_yolov8_cbam_swinblock_b200.modules as M

    torch.manual_seed(5)
    mod = M.CBAM(128).cuda()
    x = torch.randn(2, 128, 20, 20, device="cuda")
    ref = mod(x)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = mod(x.bfloat16())
    assert y.dtype == torch.bfloat16 and rel_err(y, ref) < 2e-2
    h = mod.half()
    yh = h(x.half())
    assert yh.dtype == torch.float16 and rel_err(yh, ref) < 4e-3


def test_full_size_consistency_between_modes():
    """BASELINE size (B=64, C5=256, 20x20, bf16): the fused kernel must agree with the composition of its own
    stand-alone map kernels, out = x * ca(x) * sa(x * ca(x)) (cbam.py:62-71) -- a size-independent identity."""
    import improving_yolov8_cbam_swinblock_b200.modules as M

    torch.manual_seed(0)
    mod = M.CBAM()
    mod(torch.zeros(1, 256, 2, 2))
    mod = mod.cuda()
    x = to_cl(torch.randn(64, 256, 20, 20, device="cuda").bfloat16())
    with torch.no_grad():
        out = mod(x)
        ca = mod.ca(x)
        x1 = (x.float() * ca.float())
        sa = mod.sa(to_cl(x1.bfloat16()))
        want = x1 * sa.float()
    assert ca.shape == (64, 256, 1, 1) and sa.shape == (64, 1, 20, 20)
    assert float(ca.min()) > 0 and float(ca.max()) < 1 and float(sa.min()) > 0 and float(sa.max()) < 1
    assert rel_err(out, want) < 1e-2


@pytest.mark.parametrize("shape", [(64, 256, 20, 20), (64, 128, 40, 40), (64, 64, 80, 80)])
def test_full_size_vs_oracle_on_gpu(shape):
    """BASELINE sizes (B=64; P5 = the model's use, P4 / P3 = the sweep shapes): the CUDA bf16 path against
    oracle/blocks.py (pinned by the reference-generated fixtures) evaluated in fp32 on the same GPU -- output, input
    gradient and the three weight gradients."""
    import improving_yolov8_cbam_swinblock_b200.modules as M
    from oracle import blocks as ob

    torch.manual_seed(1)
    B, C, H, W = shape
    mod = M.CBAM()
    mod(torch.zeros(1, C, 2, 2))
    mod = mod.cuda()
    x = torch.randn(shape, device="cuda").bfloat16()
    g = torch.randn(shape, device="cuda").bfloat16()
    r = mod.ca.shared_MLP[0].weight.shape[0]
    w1 = mod.ca.shared_MLP[0].weight.detach().view(r, C).clone().requires_grad_(True)
    w2 = mod.ca.shared_MLP[2].weight.detach().view(C, r).clone().requires_grad_(True)
    ws = mod.sa.conv.weight.detach().clone().requires_grad_(True)
    xo = x.float().requires_grad_(True)
    yo = ob.cbam_forward(xo, w1, w2, ws)
    yo.backward(g.float())
    xi = to_cl(x).requires_grad_(True)
    y = mod(xi)
    y.backward(to_cl(g))
    assert rel_err(y, yo) < 1e-2, rel_err(y, yo)
    assert rel_err(xi.grad, xo.grad) < 1e-2, rel_err(xi.grad, xo.grad)
    assert rel_err(mod.ca.shared_MLP[0].weight.grad.view(r, C), w1.grad) < 2e-2
    assert rel_err(mod.ca.shared_MLP[2].weight.grad.view(C, r), w2.grad) < 2e-2
    assert rel_err(mod.sa.conv.weight.grad, ws.grad) < 2e-2
