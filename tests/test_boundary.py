"""CPU: the drop-in boundary: names, ctor signatures, state_dict keys, lazy MLP + ratio rule,
deepcopy / pickle, optimizer grouping, shape-probe-only CPU behaviour, and oracle isolation."""
import copy
import inspect
import io
import os
import re

import pytest
import torch
import torch.nn as nn

import improving_yolov8_cbam_swinblock_b200 as P
from improving_yolov8_cbam_swinblock_b200.harness import graph, train
from util import ROOT, load_golden


def test_signatures_match_reference():
    assert str(inspect.signature(P.CBAM.__init__)) == "(self, channels=None)"  # cbam.py:56
    assert str(inspect.signature(P.ChannelAttention.__init__)) == "(self, in_planes=None, ratio=16)"  # cbam.py:6
    assert str(inspect.signature(P.SpatialAttention.__init__)) == "(self, kernel_size=7)"  # cbam.py:41
    # swin_block.py:24 -- the reference's parameters in order; `shift_size=0` is the trailing shifted-window extension
    # (DESIGN.md section 7-4), whose default reproduces the reference block
    assert str(inspect.signature(P.SwinBlock.__init__)) == "(self, dim, num_heads=2, window_size=7, shift_size=0)"
    assert str(inspect.signature(P.SPPF.__init__)) == "(self, c1, c2, k=5)"  # block.py:204
    with pytest.raises(AssertionError, match="3 or 7"):
        P.SpatialAttention(5)


@pytest.mark.parametrize("name,ctor", [
    ("cbam_lazy_c32", lambda: P.CBAM()), ("cbam_c64_r8", lambda: P.CBAM(64)),
    ("swin_c32_ws8_h4", lambda: P.SwinBlock(32, 4, 8)), ("sppf_module_k5", lambda: P.SPPF(16, 16, 5)),
])
def test_state_dict_keys_and_shapes_equal_reference(name, ctor):
    g = load_golden(name)
    want = {k[2:]: tuple(v.shape) for k, v in g.items() if k.startswith("w.")}
    m = ctor()
    m(torch.zeros(g["x"].shape))  # CPU zeros = the stride pass: creates CBAM's lazy MLP
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == want
    m.load_state_dict({k[2:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("w.")})


def test_lazy_mlp_and_ratio_rule():
    m = P.CBAM()
    assert m.ca.shared_MLP is None and m.ca.ratio == 16  # cbam.py:59 with channels=None
    y = m(torch.zeros(1, 576, 4, 4))
    assert y.shape == (1, 576, 4, 4) and m.ca.shared_MLP[0].weight.shape == (36, 576, 1, 1)
    assert P.CBAM(64).ca.shared_MLP[0].weight.shape == (8, 64, 1, 1)  # ratio 8 below 128 channels
    assert P.CBAM(8).ca.shared_MLP[0].weight.shape == (1, 8, 1, 1)  # max(1, .)
    assert P.ChannelAttention(32)(torch.zeros(2, 32, 5, 5)).shape == (2, 32, 1, 1)
    assert P.SpatialAttention()(torch.zeros(2, 32, 5, 5)).shape == (2, 1, 5, 5)


def test_cpu_is_shape_probe_only():
    m = P.SwinBlock(16)
    assert m(torch.zeros(1, 16, 9, 9)).shape == (1, 16, 9, 9)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.randn(1, 16, 9, 9))
    with pytest.raises(RuntimeError, match="no CPU path"):
        P.CBAM(16)(torch.ones(1, 16, 4, 4))


def test_graph_builds_like_parse_model():
    for scale, (c4, c5, nparam) in {"n": (128, 256, 3726642), "s": (256, 512, None)}.items():
        m = graph.DetectionGraph(P.BLOCKS, scale, 80)
        assert isinstance(m.model[7], P.SwinBlock) and m.model[7].dim == c4 and isinstance(m.model[16], P.SwinBlock)
        assert isinstance(m.model[10], P.CBAM) and m.model[10].ca.shared_MLP[0].in_channels == c5
        assert isinstance(m.model[11], P.SPPF) and m.model[11].k == 5 and m.model[12].k == 7
        assert m.model[11].cv1.conv.out_channels == c5 // 2 and m.model[11].cv2.conv.in_channels == 2 * c5
        assert m.stride.tolist() == [8.0, 16.0, 32.0]
        if nparam:
            assert sum(p.numel() for p in m.parameters()) == nparam  # == reference DetectionModel (probe)


def test_deepcopy_pickle_half_and_param_groups():
    m = graph.DetectionGraph(P.BLOCKS, "n", 3)
    m2 = copy.deepcopy(m)
    buf = io.BytesIO()
    torch.save(m, buf)
    buf.seek(0)
    m3 = torch.load(buf, weights_only=False)
    assert list(m2.state_dict()) == list(m.state_dict()) == list(m3.state_dict())
    assert all(p.dtype == torch.float16 for p in copy.deepcopy(m).half().parameters())
    g0, g1, g2 = (g["params"] for g in train.param_groups(m.model[7]))
    ids = lambda ps: {id(p) for p in ps}  # noqa: E731
    sw = m.model[7]
    assert ids(g0) == ids([sw.attn.in_proj_weight, sw.attn.out_proj.weight, sw.mlp[0].weight, sw.mlp[2].weight])
    assert ids(g1) == ids([sw.norm1.weight, sw.norm2.weight])
    assert len(g2) == 6  # norm biases, in_proj_bias, out_proj.bias, mlp biases (trainer.py:818-827)
    for mod in m.modules():  # nothing but tensors / numbers on the modules: no ctypes handles
        assert not any("ctypes" in type(v).__module__ for v in vars(mod).values())


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "improving_yolov8_cbam_swinblock_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f"{f} imports the oracle"
                assert "/root/reference" not in txt, f"{f} reads the reference at run time"
