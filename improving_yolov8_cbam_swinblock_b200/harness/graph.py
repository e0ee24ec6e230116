"""Caller-side model graph for the throughput harness: the stock (non-hot-path) YOLOv8 pieces.

BASELINE.json's metric is whole-model training images/s of the fork's custom YOLOv8-CBAM-Swin yaml
(``ultralytics/cfg/models/v8/yolov8.yaml:734-776``).  The reference package cannot travel to the GPU box, so
the harness carries a compact plain-PyTorch statement of the *callers* of the hot path -- ``Conv``
(``nn/modules/conv.py:37-91``), ``Bottleneck``/``C2f`` (``nn/modules/block.py:279-310,340-360``), ``Concat``
(``conv.py:655``), ``DFL``/``Detect`` (``block.py:58-77``, ``head.py:23-84``) and a yaml-row parser with
``parse_model``'s channel rules (``nn/tasks.py:1340-1517``).  These stay stock cuDNN/cuBLAS modules (SURVEY
section 2.1: "OUT OF SCOPE -- stock path, part of e2e time but not rewritten"); the three hot-path blocks
(CBAM / SwinBlock / SPPF) are resolved *by name* from the ``blocks`` table the caller passes in, which is
the same contract as ``globals()[m]`` in ``tasks.py:1438``.  ``state_dict`` keys equal the reference
``DetectionModel``'s (``model.<i>.<...>``); tests/test_harness_vs_reference.py checks that and the outputs.
"""
from __future__ import annotations

import copy
import math

import torch
import torch.nn as nn

# yolov8.yaml:734-776 (the ONE active architecture of the fork), transcribed as data.
# The two SwinBlock rows carry [256] in the file and are not width-scaled by parse_model
# (tasks.py:1503-1504), so the yaml only builds at scale "s"; per the author's note at yolov8.yaml:664
# the n/s/m/l variants use dim 128/256/384/512.  SWIN_DIM below applies that.
SCALES = {"n": (0.33, 0.25, 1024), "s": (0.33, 0.50, 1024), "m": (0.67, 0.75, 768),
          "l": (1.00, 1.00, 512), "x": (1.00, 1.25, 512)}
SWIN_DIM = {"n": 128, "s": 256, "m": 384, "l": 512, "x": 640}
BACKBONE = [
    [-1, 1, "Conv", [64, 3, 2]],
    [-1, 1, "Conv", [128, 3, 2]],
    [-1, 3, "C2f", [128, True]],
    [-1, 1, "Conv", [256, 3, 2]],
    [-1, 6, "C2f", [256, True]],
    [-1, 1, "Conv", [512, 3, 2]],
    [-1, 6, "C2f", [512, True]],
    [-1, 1, "SwinBlock", [256]],
    [-1, 1, "Conv", [1024, 3, 2]],
    [-1, 3, "C2f", [1024, True]],
    [-1, 1, "CBAM", []],
    [-1, 1, "SPPF", [1024, 5]],
    [-1, 1, "SPPF", [1024, 7]],
]
HEAD = [
    [-1, 1, "nn.Upsample", [None, 2, "nearest"]],
    [[-1, 7], 1, "Concat", [1]],
    [-1, 3, "C2f", [512]],
    [-1, 1, "SwinBlock", [256]],
    [-1, 1, "nn.Upsample", [None, 2, "nearest"]],
    [[-1, 4], 1, "Concat", [1]],
    [-1, 3, "C2f", [256]],
    [-1, 1, "Conv", [256, 3, 2]],
    [[-1, 16], 1, "Concat", [1]],
    [-1, 3, "C2f", [512]],
    [-1, 1, "Conv", [512, 3, 2]],
    [[-1, 10], 1, "Concat", [1]],
    [-1, 3, "C2f", [1024]],
    [[19, 22, 25], 1, "Detect", ["nc"]],
]


def model_dict(scale: str = "n", nc: int = 80) -> dict:
    """The active yaml as a dict, SwinBlock dims set for ``scale``."""
    d = {"nc": nc, "scale": scale, "scales": copy.deepcopy(SCALES),
         "backbone": copy.deepcopy(BACKBONE), "head": copy.deepcopy(HEAD)}
    d["backbone"][7][3] = [SWIN_DIM[scale]]
    d["head"][3][3] = [SWIN_DIM[scale]]
    return d


def make_divisible(x, divisor):
    return math.ceil(x / divisor) * divisor


class Conv(nn.Module):
    """conv(bias=False) -> BatchNorm2d -> SiLU (conv.py:37-91)."""

    def __init__(self, c1, c2, k=1, s=1, p=None, g=1, d=1, act=True):
        super().__init__()
        if p is None:
            p = k // 2 if isinstance(k, int) else [x // 2 for x in k]
        self.conv = nn.Conv2d(c1, c2, k, s, p, groups=g, dilation=d, bias=False)
        self.bn = nn.BatchNorm2d(c2)
        self.act = nn.SiLU() if act is True else act if isinstance(act, nn.Module) else nn.Identity()

    epilogue = None  # DetectionGraph sets blocks["conv_epilogue"] here: callable (conv_module, y) -> act(bn(y))
    conv_fn = None   # per instance: callable (nn.Conv2d, x) -> conv(x); set on the first layer from blocks["stem_conv"]
    w16 = None       # per instance, set by the Trainer: a 16-bit LEAF copy of conv.weight, refreshed once per step for all layers
                     # together; its gradient is copied back the same way (instead of autocast's cast kernel per layer and direction)

    def forward(self, x):
        if self.w16 is not None and self.conv_fn is None and self.training and x.dtype == self.w16.dtype and torch.is_grad_enabled():
            c = self.conv
            y = nn.functional.conv2d(x, self.w16, None, c.stride, c.padding, c.dilation, c.groups)
            return self.act(self.bn(y)) if self.epilogue is None else self.epilogue(self, y)
        y = self.conv(x) if self.conv_fn is None else self.conv_fn(self.conv, x)
        return self.act(self.bn(y)) if self.epilogue is None else self.epilogue(self, y)


class Bottleneck(nn.Module):
    def __init__(self, c1, c2, shortcut=True, g=1, k=(3, 3), e=0.5):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, k[0], 1)
        self.cv2 = Conv(c_, c2, k[1], 1, g=g)
        self.add = shortcut and c1 == c2

    fork = None   # blocks["fork"] (see C2f.fork): with a shortcut, x feeds cv1 AND the residual add

    def forward(self, x):
        if self.add and self.fork is not None:
            # the add's gradient can be a channel slice of the C2f concat's gradient (last bottleneck): summed with cv1's input
            # gradient by the fan-in kernel instead of ATen's generic strided add
            xa, xb = self.fork(x, 2)
            return xa + self.cv2(self.cv1(xb))
        y = self.cv2(self.cv1(x))
        return x + y if self.add else y


class C2f(nn.Module):
    def __init__(self, c1, c2, n=1, shortcut=False, g=1, e=0.5):
        super().__init__()
        self.c = int(c2 * e)
        self.cv1 = Conv(c1, 2 * self.c, 1, 1)
        self.cv2 = Conv((2 + n) * self.c, c2, 1)
        self.m = nn.ModuleList(Bottleneck(self.c, self.c, shortcut, g, k=(3, 3), e=1.0) for _ in range(n))

    # DetectionGraph sets blocks["concat"] / blocks["chunk"] here (the channel concat / chunk seams);
    # None = torch.cat / Tensor.chunk
    concat = None
    chunk = None
    fork = None   # blocks["fork"]: (x, n) -> n aliases of x whose gradients are summed by ONE fan-in kernel (None = autograd's adds)

    def forward(self, x):
        t = self.cv1(x)
        y = list(t.chunk(2, 1) if self.chunk is None else self.chunk(t, 2))
        # the bottlenecks' 3x3 convs and shortcut adds want a dense operand: one vectorised copy of the second half here
        # instead of cuDNN's generic strided `.contiguous()` and a strided element-wise add per bottleneck
        if self.fork is not None:   # y[1] feeds the concat and the first bottleneck
            y[-1], inp = self.fork(y[-1], 2)
        else:
            inp = y[-1]
        inp = inp if self.concat is None else self.concat([inp])
        last = len(self.m) - 1
        for k, m in enumerate(self.m):
            inp = m(inp)
            if self.fork is not None and k < last:   # a bottleneck output feeds the concat and the next bottleneck
                a, inp = self.fork(inp, 2)
                y.append(a)
            else:
                y.append(inp)
        return self.cv2(torch.cat(y, 1) if self.concat is None else self.concat(y))


class Upsample(nn.Upsample):
    """``nn.Upsample`` (yolov8.yaml head rows) that keeps the activation dtype under autocast.  torch's autocast runs
    ``upsample_nearest2d`` in fp32, which turns the following ``Concat`` and its consumers' input casts into fp32 passes;
    nearest-neighbour upsampling only copies values, so staying in bf16/f16 is bit-identical (dtype/layout at
    the seams of the path).  Same constructor, no parameters, so state_dicts are unaffected."""

    upsample = None  # DetectionGraph sets blocks["upsample"] here: callable (x, sh, sw) for integer nearest factors

    def forward(self, x):
        sf = self.scale_factor
        if self.upsample is not None and x.is_cuda and self.mode == "nearest" and self.size is None and sf is not None:
            sh, sw = (sf, sf) if not isinstance(sf, (tuple, list)) else sf
            if float(sh).is_integer() and float(sw).is_integer():
                return self.upsample(x, int(sh), int(sw))   # dtype-preserving, like the branch below
        if x.is_cuda and self.mode == "nearest" and torch.is_autocast_enabled("cuda"):
            with torch.autocast("cuda", enabled=False):
                return super().forward(x)
        return super().forward(x)


class Concat(nn.Module):
    def __init__(self, dimension=1):
        super().__init__()
        self.d = dimension

    concat = None

    def forward(self, x):
        return torch.cat(x, self.d) if self.concat is None or self.d != 1 else self.concat(x)


class DFL(nn.Module):
    def __init__(self, c1=16):
        super().__init__()
        self.conv = nn.Conv2d(c1, 1, 1, bias=False).requires_grad_(False)
        self.conv.weight.data[:] = torch.arange(c1, dtype=torch.float).view(1, c1, 1, 1)
        self.c1 = c1

    def forward(self, x):
        b, _, a = x.shape
        return self.conv(x.view(b, 4, self.c1, a).transpose(2, 1).softmax(1)).view(b, 4, a)


def make_anchors(feats, strides, grid_cell_offset=0.5):
    """Anchor centres + stride per anchor for the three maps (utils/tal.py:330-343)."""
    pts, st = [], []
    for f, s in zip(feats, strides):
        h, w = f.shape[2:]
        sx = torch.arange(w, device=f.device, dtype=f.dtype) + grid_cell_offset
        sy = torch.arange(h, device=f.device, dtype=f.dtype) + grid_cell_offset
        sy, sx = torch.meshgrid(sy, sx, indexing="ij")
        pts.append(torch.stack((sx, sy), -1).view(-1, 2))
        st.append(torch.full((h * w, 1), float(s), dtype=f.dtype, device=f.device))
    return torch.cat(pts), torch.cat(st)


class Detect(nn.Module):
    """Legacy (v8) Detect head (head.py:23-84): training mode returns the three raw maps."""

    def __init__(self, nc=80, ch=()):
        super().__init__()
        self.nc, self.nl, self.reg_max = nc, len(ch), 16
        self.no = nc + self.reg_max * 4
        self.stride = torch.zeros(self.nl)
        c2, c3 = max((16, ch[0] // 4, self.reg_max * 4)), max(ch[0], min(self.nc, 100))
        self.cv2 = nn.ModuleList(
            nn.Sequential(Conv(x, c2, 3), Conv(c2, c2, 3), nn.Conv2d(c2, 4 * self.reg_max, 1)) for x in ch)
        self.cv3 = nn.ModuleList(
            nn.Sequential(Conv(x, c3, 3), Conv(c3, c3, 3), nn.Conv2d(c3, self.nc, 1)) for x in ch)
        self.dfl = DFL(self.reg_max)

    concat = None
    head_conv = None        # the last 1x1 convolution (+ bias) of every branch as a fused GEMM (None = nn.Conv2d)
    split_outputs = False   # training: hand the loss the (box, cls) maps of each level separately (no channel concat: the fused
                            # classification-loss kernel reads the class maps in place and the gradients stay dense)

    def _branch(self, seq, x):
        if self.head_conv is None:
            return seq(x)
        return self.head_conv(seq[2], seq[1](seq[0](x)))

    def forward(self, x):
        cat = (lambda ts: torch.cat(ts, 1)) if self.concat is None else self.concat
        if self.training and self.split_outputs:
            return [(self._branch(self.cv2[i], x[i]), self._branch(self.cv3[i], x[i])) for i in range(self.nl)]
        x = [cat((self._branch(self.cv2[i], x[i]), self._branch(self.cv3[i], x[i]))) for i in range(self.nl)]
        if self.training:
            return x
        shape = x[0].shape
        x_cat = torch.cat([xi.reshape(shape[0], self.no, -1) for xi in x], 2)
        anchors, strides = (t.transpose(0, 1) for t in make_anchors(x, self.stride, 0.5))
        box, cls = x_cat.split((self.reg_max * 4, self.nc), 1)
        lt, rb = self.dfl(box).chunk(2, 1)
        x1y1, x2y2 = anchors.unsqueeze(0) - lt, anchors.unsqueeze(0) + rb
        dbox = torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), 1) * strides
        return torch.cat((dbox, cls.sigmoid()), 1), x

    def bias_init(self):
        for a, b, s in zip(self.cv2, self.cv3, self.stride):
            a[-1].bias.data[:] = 1.0
            b[-1].bias.data[: self.nc] = math.log(5 / self.nc / (640 / s) ** 2)


_BASE = {"Conv", "C2f", "SPPF"}  # rows whose args get [c1, width-scaled c2, ...] (tasks.py:1375-1412,1445-1453)


class DetectionGraph(nn.Module):
    """yaml rows -> nn.Sequential + the ``_predict_once`` loop (tasks.py:152-179,321-369).

    ``blocks``: name -> class for {"CBAM", "SwinBlock", "SPPF"}; this is the drop-in boundary (the same lookup
    ``parse_model`` does with ``globals()[m]``, tasks.py:1438).
    """

    def __init__(self, blocks: dict, scale="n", nc=80, ch=3, stride_probe=256):
        super().__init__()
        d = model_dict(scale, nc)
        depth, width, max_ch = d["scales"][scale]
        table = {"Conv": Conv, "C2f": C2f, "Concat": Concat, "Detect": Detect, **blocks}
        chans, layers, save = [ch], [], []
        for i, (f, n, m, args) in enumerate(d["backbone"] + d["head"]):
            args = [nc if a == "nc" else a for a in args]
            n = max(round(n * depth), 1) if n > 1 else n
            if m in _BASE:
                c1, c2 = chans[f], args[0]
                if c2 != nc:
                    c2 = make_divisible(min(c2, max_ch) * width, 8)
                args = [c1, c2, *args[1:]]
                if m == "C2f":
                    args.insert(2, n)
                    n = 1
            elif m == "Concat":
                c2 = sum(chans[x] for x in f)
            elif m == "Detect":
                args.append([chans[x] for x in f])
                c2 = None
            else:  # CBAM / SwinBlock / nn.Upsample: args verbatim, c2 = ch[f]  (tasks.py:1503-1504)
                c2 = chans[f]
            cls = Upsample if m == "nn.Upsample" else getattr(nn, m[3:]) if m.startswith("nn.") else table[m]
            mod = nn.Sequential(*(cls(*args) for _ in range(n))) if n > 1 else cls(*args)
            mod.i, mod.f, mod.type = i, f, m
            save.extend(x % i for x in ([f] if isinstance(f, int) else f) if x != -1)
            layers.append(mod)
            if i == 0:
                chans = []
            chans.append(c2)
        self.model = nn.Sequential(*layers)
        epi = blocks.get("conv_epilogue")  # fused BN+SiLU epilogue for the Conv callers (None = stock ops)
        if epi is not None:
            for m_ in self.modules():
                if isinstance(m_, Conv):
                    m_.epilogue = epi
        cat_, chunk_ = blocks.get("concat"), blocks.get("chunk")  # channel concat / chunk at the seams
        for m_ in self.modules():
            if cat_ is not None and isinstance(m_, (C2f, Concat, Detect)):
                m_.concat = cat_
            if chunk_ is not None and isinstance(m_, C2f):
                m_.chunk = chunk_
            if blocks.get("fork") is not None and isinstance(m_, (C2f, Bottleneck)):
                m_.fork = blocks["fork"]
            if blocks.get("upsample") is not None and isinstance(m_, Upsample):
                m_.upsample = blocks["upsample"]
            if blocks.get("head_conv") is not None and isinstance(m_, Detect):
                m_.head_conv = blocks["head_conv"]
        if blocks.get("conv3x3") is not None:   # narrow 3x3 convolutions: weight gradient on the hand-written kernel
            for m_ in self.modules():
                if isinstance(m_, Conv) and m_.conv.kernel_size == (3, 3) and m_.conv.in_channels == 16 \
                        and m_.conv.out_channels in (16, 32):
                    m_.conv_fn = blocks["conv3x3"]
        if blocks.get("stem_conv") is not None and isinstance(self.model[0], Conv):   # the 3-channel first layer (yaml backbone row 0)
            self.model[0].conv_fn = blocks["stem_conv"]
        self.save = sorted(save)
        # consumers per layer output (the next layer when its `from` is -1, plus every later row that names it): a saved layer
        # with several consumers hands each its own alias, so that their gradients meet in one fan-in kernel
        self._fork = blocks.get("fork")
        self._consumers = [0] * len(layers)
        for mod in layers:
            for j in ([mod.f] if isinstance(mod.f, int) else mod.f):
                src = mod.i - 1 if j == -1 else j % mod.i
                if src >= 0:
                    self._consumers[src] += 1
        self.nc, self.scale = nc, scale
        det = self.model[-1]
        # stride pass on CPU zeros (tasks.py:350-364); CBAM creates its lazy MLP here (cbam.py:31-33)
        self.train()
        with torch.no_grad():
            outs = self._predict_once(torch.zeros(1, ch, stride_probe, stride_probe))
        det.stride = torch.tensor([stride_probe / o.shape[-2] for o in outs])
        self.stride = det.stride
        det.bias_init()
        for mm in self.modules():  # initialize_weights (torch_utils.py:462-472)
            if isinstance(mm, nn.BatchNorm2d):
                mm.eps, mm.momentum = 1e-3, 0.03

    def _predict_once(self, x):
        y = []   # per layer: the aliases of its output not yet handed to a consumer (tasks.py:171-176 keeps `y[j]` itself)
        for m in self.model:
            if m.i > 0:
                src = [m.f] if isinstance(m.f, int) else m.f
                ins = [y[m.i - 1 if j == -1 else j].pop() for j in src]
                x = ins[0] if isinstance(m.f, int) else ins
            x = m(x)
            n = self._consumers[m.i]
            if self._fork is not None and n > 1 and torch.is_tensor(x):
                y.append(list(self._fork(x, n)))
            else:
                y.append([x] * max(n, 1))
        return x

    def forward(self, x):
        return self._predict_once(x)
