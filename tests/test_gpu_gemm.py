"""-m gpu: the hand-written tcgen05 GEMM (b200_gemm_nt) vs torch.matmul in fp32 on the same bf16/f16 inputs."""
import pytest
import torch

from util import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 128, 128), (1000, 384, 128), (112896, 128, 512), (4097, 512, 128),
                                   (77, 64, 64), (3000, 192, 192), (640, 256, 256)])
def test_gemm_bias(dtype, M, N, K):
    from improving_yolov8_cbam_swinblock_b200 import gemm_tc

    torch.manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda").to(dtype)
    w = (torch.randn(N, K, device="cuda") / K ** 0.5)
    b = torch.randn(N, device="cuda")
    assert gemm_tc.supports(a, N, K)
    got = gemm_tc.gemm_nt(a, w, b)
    want = a.float() @ w.to(dtype).float().t() + b
    assert got.shape == (M, N) and got.dtype == dtype
    assert rel_err(got, want) < (4e-3 if dtype == torch.bfloat16 else 6e-4), rel_err(got, want)
    # element-wise: every output within one rounding step of the fp32 result
    torch.testing.assert_close(got.float(), want, rtol=2e-2 if dtype == torch.bfloat16 else 3e-3, atol=2e-2)


@pytest.mark.parametrize("M,N,K", [(1024, 512, 128), (5000, 128, 512)])
def test_gemm_epilogues(M, N, K):
    from improving_yolov8_cbam_swinblock_b200 import gemm_tc

    torch.manual_seed(1)
    dt = torch.bfloat16
    a = torch.randn(M, K, device="cuda").to(dt)
    w = torch.randn(N, K, device="cuda") / K ** 0.5
    b = torch.randn(N, device="cuda")
    r = torch.randn(M, N, device="cuda").to(dt)
    pre = a.float() @ w.to(dt).float().t() + b
    h, hp = gemm_tc.gemm_nt(a, w, b, gemm_tc.EPI_BIAS_GELU, want_preact=True)
    assert rel_err(hp, pre) < 4e-3
    assert torch.equal(h, torch.nn.functional.gelu(hp.float()).to(dt)) or rel_err(h, torch.nn.functional.gelu(hp.float())) < 4e-3
    y = gemm_tc.gemm_nt(a, w, b, gemm_tc.EPI_BIAS_RES, residual=r)
    assert rel_err(y, pre + r.float()) < 4e-3


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,N,K", [(1024, 512, 128), (5000, 128, 512), (77, 64, 64)])
def test_gemm_epilogue_mul_gelugrad(dtype, M, N, K):
    """EPI_MUL_GELUGRAD (epilogue 3): D = (A W^T) * gelu'(R), the GELU backward fused into the mlp.2 data-gradient GEMM
    (swin_block.py:53 backward), against erf-GELU's analytic derivative in fp64."""
    from improving_yolov8_cbam_swinblock_b200 import gemm_tc

    torch.manual_seed(M + N)
    a = torch.randn(M, K, device="cuda").to(dtype)
    w = torch.randn(N, K, device="cuda") / K ** 0.5
    r = (2.5 * torch.randn(M, N, device="cuda")).to(dtype)
    got = gemm_tc.gemm_nt(a, w, None, gemm_tc.EPI_MUL_GELUGRAD, residual=r)
    rd = r.double()
    dgelu = 0.5 * (1 + torch.erf(rd / 2 ** 0.5)) + rd * torch.exp(-0.5 * rd * rd) / (2 * torch.pi) ** 0.5
    want = (a.double() @ w.to(dtype).double().t()) * dgelu
    assert got.dtype == dtype and rel_err(got, want) < (4e-3 if dtype == torch.bfloat16 else 8e-4), rel_err(got, want)


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 128, 256), (384, 128, 4096), (128, 512, 112896), (72, 200, 1000)])
def test_gemm_splitk_all_majors(a_mn, b_mn, M, N, K):
    """K-major and MN-major operand descriptors: D = A B^T in f32, A/B given in either storage order."""
    from improving_yolov8_cbam_swinblock_b200 import gemm_tc

    torch.manual_seed(M + N + (K % 1000))
    dt = torch.bfloat16
    A = torch.randn(M, K, device="cuda").to(dt)
    Bm = torch.randn(N, K, device="cuda").to(dt)
    a = A.t().contiguous() if a_mn else A
    b = Bm.t().contiguous() if b_mn else Bm
    got, cs = gemm_tc.gemm_splitk(a, b, a_mn, b_mn, want_colsum=True)
    want = A.double() @ Bm.double().t()
    assert rel_err(got, want) < 1e-5, rel_err(got, want)
    assert rel_err(cs, A.double().sum(1)) < 1e-5, rel_err(cs, A.double().sum(1))  # fused bias-gradient column sums


@pytest.mark.parametrize("dtype,tol", [(torch.bfloat16, 1e-2), (torch.float16, 2e-3)])
@pytest.mark.parametrize("shape,c2", [((2, 64, 20, 20), 64), ((3, 80, 40, 40), 80), ((2, 80, 20, 20), 80), ((1, 64, 80, 80), 64),
                                      ((2, 96, 12, 10), 144)])
def test_head_conv_with_bias_matches_conv2d(dtype, tol, shape, c2):
    """the Detect head's last 1x1 convolution (+ bias, head.py:45-62) as b200_gemm_nt (bias in the epilogue) with
    dW and db out of one b200_gemm_splitk pass; truth = F.conv2d in fp32 on the same (rounded) inputs."""
    from improving_yolov8_cbam_swinblock_b200 import functional as Fb

    torch.manual_seed(sum(shape) + c2)
    conv = torch.nn.Conv2d(shape[1], c2, 1).cuda()
    with torch.no_grad():
        conv.bias.normal_()
    x = torch.randn(shape, device="cuda").to(dtype).contiguous(memory_format=torch.channels_last)
    g = torch.randn(shape[0], c2, *shape[2:], device="cuda").to(dtype).contiguous(memory_format=torch.channels_last)
    assert Fb.conv1x1_supported(x, conv, allow_bias=True) and not Fb.conv1x1_supported(x, conv)
    from improving_yolov8_cbam_swinblock_b200 import _lib

    _launches = _lib.launch_count
    launches = _launches()
    xg = x.clone().requires_grad_(True)
    y = Fb.head_conv(conv, xg)
    y.backward(g)
    assert _launches() - launches >= 3, "head_conv did not run the hand-written GEMMs"
    got = (y.detach().float(), xg.grad.float(), conv.weight.grad.clone(), conv.bias.grad.clone())
    conv.zero_grad()
    w32 = conv.weight.detach().to(dtype).float().requires_grad_(True)   # the GEMM consumes the rounded weight
    b32 = conv.bias.detach().clone().requires_grad_(True)
    x32 = x.float().requires_grad_(True)
    y32 = torch.nn.functional.conv2d(x32, w32, b32)
    y32.backward(g.float())
    for name, a, b in zip(("y", "gx", "gw", "gb"), got, (y32.detach(), x32.grad, w32.grad, b32.grad)):
        assert rel_err(a, b) <= tol, f"{name}: {rel_err(a, b):.2e}"
    # under autocast the f32 feature map is cast first, exactly as F.conv2d would
    with torch.autocast("cuda", dtype=dtype):
        y2 = Fb.head_conv(conv, x.float())
    assert y2.dtype == dtype and rel_err(y2.float(), y32.detach()) <= tol
