"""-m gpu: fused BatchNorm2d(+SiLU) Conv epilogue (csrc/conv_epilogue.cu) vs the reference-generated
fixtures, the fp64 oracle and torch's own BatchNorm2d + SiLU.  Bars: fp32 rtol 1e-5; bf16/f16 <= 2e-2 / 4e-3 relative
error vs the fp64 oracle for the output and every gradient; running statistics to 1e-5."""
import pytest
import torch
import torch.nn as nn

from util import assert_close_f32, load_golden, rel_err, to_cl

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["conv_k1_c16", "conv_k3_c32"])
def test_golden_fp32(name):
    from improving_yolov8_cbam_swinblock_b200 import functional as Fb

    g = load_golden(name)
    C = g["w0.bn.weight"].shape[0]
    bn = nn.BatchNorm2d(C, eps=float(g["bn_eps"]), momentum=float(g["bn_momentum"]))
    bn.load_state_dict({k[6:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("w0.bn.")})
    bn = bn.cuda().train()
    x = to_cl(torch.from_numpy(g["conv_out"]).cuda()).requires_grad_(True)
    assert Fb.bn_act_supported(x, bn)
    z = Fb.bn_act(x, bn, True)
    z.backward(torch.from_numpy(g["gy"]).cuda())
    assert_close_f32(z, torch.from_numpy(g["y"]), name + " y")
    assert_close_f32(bn.weight.grad, torch.from_numpy(g["gw.bn.weight"]), name + " g_gamma", rtol=1e-4, atol=1e-5)
    assert_close_f32(bn.bias.grad, torch.from_numpy(g["gw.bn.bias"]), name + " g_beta", rtol=1e-4, atol=1e-5)
    assert_close_f32(bn.running_mean, torch.from_numpy(g["w.bn.running_mean"]), name + " running_mean")
    assert_close_f32(bn.running_var, torch.from_numpy(g["w.bn.running_var"]), name + " running_var")
    assert int(bn.num_batches_tracked) == int(g["w.bn.num_batches_tracked"])


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 3e-6), (torch.bfloat16, 2e-2), (torch.float16, 4e-3)])
@pytest.mark.parametrize("shape", [(4, 16, 32, 32), (2, 80, 20, 20), (8, 256, 20, 20), (3, 64, 17, 13), (2, 32, 96, 96)])
@pytest.mark.parametrize("silu", [True, False])
def test_vs_oracle_train(dtype, tol, shape, silu):
    from improving_yolov8_cbam_swinblock_b200 import functional as Fb
    from oracle import blocks as ob

    torch.manual_seed(sum(shape))
    B, C, H, W = shape
    bn = nn.BatchNorm2d(C, eps=1e-3, momentum=0.03)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_(0, 0.3)
        bn.running_mean.normal_(0, 0.1)
        bn.running_var.uniform_(0.5, 2.0)
    x = (torch.randn(shape) * 1.7 + 0.8).to(dtype)     # non-zero mean: the shifted-sum path matters
    gz = torch.randn(shape).to(dtype)
    xo = x.double().requires_grad_(True)
    go, bo = bn.weight.detach().double().requires_grad_(True), bn.bias.detach().double().requires_grad_(True)
    zo, rm, rv = ob.bn_act_forward(xo, go, bo, bn.running_mean.double(), bn.running_var.double(), True, 0.03, 1e-3, silu)
    zo.backward(gz.double())
    bn = bn.cuda().train()
    xc = to_cl(x.cuda()).requires_grad_(True)
    z = Fb.bn_act(xc, bn, silu)
    assert z.dtype == dtype and z.is_contiguous(memory_format=torch.channels_last)
    z.backward(to_cl(gz.cuda()))
    assert rel_err(z.cpu(), zo) <= tol
    assert rel_err(xc.grad.cpu(), xo.grad) <= max(tol, 2e-5)
    assert rel_err(bn.weight.grad.cpu(), go.grad) <= max(tol, 2e-5)
    assert rel_err(bn.bias.grad.cpu(), bo.grad) <= max(tol, 2e-5)
    assert rel_err(bn.running_mean.cpu(), rm) <= max(tol * 0.1, 1e-5)
    assert rel_err(bn.running_var.cpu(), rv) <= max(tol * 0.1, 1e-5)


def test_eval_mode_and_determinism_and_stock_equivalence():
    from improving_yolov8_cbam_swinblock_b200 import functional as Fb

    torch.manual_seed(7)
    bn = nn.BatchNorm2d(64, eps=1e-3, momentum=0.03).cuda()
    with torch.no_grad():
        bn.running_mean.normal_(0, 0.2)
        bn.running_var.uniform_(0.5, 2.0)
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_(0, 0.2)
    x = to_cl(torch.randn(4, 64, 24, 24, device="cuda"))
    g = to_cl(torch.randn(4, 64, 24, 24, device="cuda"))
    bn.eval()
    xa = x.clone().requires_grad_(True)
    za = Fb.bn_act(xa, bn, True)
    za.backward(g)
    ga, gb = bn.weight.grad.clone(), bn.bias.grad.clone()
    bn.zero_grad()
    xb = x.clone().requires_grad_(True)
    zb = torch.nn.functional.silu(bn(xb))
    zb.backward(g)
    assert_close_f32(za, zb, "eval z")
    assert_close_f32(xa.grad, xb.grad, "eval gx", rtol=1e-4, atol=1e-5)
    assert_close_f32(ga, bn.weight.grad, "eval g_gamma", rtol=1e-4, atol=1e-5)
    assert_close_f32(gb, bn.bias.grad, "eval g_beta", rtol=1e-4, atol=1e-5)
    # training mode: run-to-run bit-identical (fixed-order merges, no atomics)
    bn.train()
    outs = []
    for _ in range(2):
        bn2 = nn.BatchNorm2d(64, eps=1e-3, momentum=0.03).cuda().train()
        xx = x.clone().requires_grad_(True)
        zz = Fb.bn_act(xx, bn2, True)
        zz.backward(g)
        outs.append((zz.detach().clone(), xx.grad.clone(), bn2.weight.grad.clone(), bn2.running_var.clone()))
    for u, v in zip(*outs):
        assert torch.equal(u, v)


def test_full_size_properties():
    """BASELINE size (P1 stem output: B=64, C=16, 320x320, bf16 = 105 M elements): size-independent identities --
    normalised output statistics, and shift/scale invariance of BatchNorm (bn(a*x+b) == bn(x) for a > 0)."""
    from improving_yolov8_cbam_swinblock_b200 import functional as Fb

    torch.manual_seed(0)
    bn = nn.BatchNorm2d(16, eps=1e-3, momentum=0.03).cuda().train()
    x = to_cl((torch.randn(64, 16, 320, 320, device="cuda") * 2 + 1).bfloat16())
    with torch.no_grad():
        y = Fb.bn_act(x, bn, False).float()
        m = y.mean(dim=(0, 2, 3))
        v = y.var(dim=(0, 2, 3), unbiased=False)
        assert float(m.abs().max()) < 5e-3 and float((v - 1).abs().max()) < 2e-2
        bn2 = nn.BatchNorm2d(16, eps=1e-3, momentum=0.03).cuda().train()
        y2 = Fb.bn_act(to_cl((x.float() * 4 + 8).bfloat16()), bn2, False).float()   # exact in bf16 (power-of-two scale)
        assert rel_err(y2, y) < 1e-2


def test_graph_with_and_without_fused_epilogue():
    """The harness graph gives the same loss / gradients with the fused epilogue as with the stock BatchNorm2d + SiLU."""
    import improving_yolov8_cbam_swinblock_b200 as P
    from improving_yolov8_cbam_swinblock_b200.harness import graph

    torch.backends.cudnn.allow_tf32 = False  # stock convs would otherwise run TF32 and drown the comparison
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(1)
    blocks_stock = {k: v for k, v in P.BLOCKS.items() if k != "conv_epilogue"}
    a = graph.DetectionGraph(P.BLOCKS, "n", 8).cuda().to(memory_format=torch.channels_last).train()
    b = graph.DetectionGraph(blocks_stock, "n", 8).cuda().to(memory_format=torch.channels_last).train()
    b.load_state_dict(a.state_dict())
    x = to_cl(torch.rand(2, 3, 128, 128, device="cuda"))
    fa, fb = a(x), b(x)
    for u, v in zip(fa, fb):
        assert rel_err(u, v) < 1e-4
    sum(f.square().mean() for f in fa).backward()
    sum(f.square().mean() for f in fb).backward()
    gmax = max(float(p.grad.abs().max()) for p in b.parameters() if p.grad is not None)
    for (n, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
        if pb.grad is None:
            continue
        # parameters in front of a normalisation (e.g. a bias) have exact-zero gradients: compare on an absolute scale
        err = float((pa.grad - pb.grad).abs().max())
        assert err <= 5e-3 * max(float(pb.grad.abs().max()), 1e-3 * gmax), (n, err)
    for (ka, va), (_, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        if "running" in ka or "num_batches" in ka:
            assert rel_err(va.float(), vb.float()) < 5e-3, ka
