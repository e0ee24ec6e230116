"""TEST INFRASTRUCTURE ONLY -- one training step of the UNMODIFIED reference (``oracle/_ref`` or ``/root/reference``).

Drives the reference's own classes through the hot loop of ``engine/trainer.py:367-399,614-622`` without its data
pipeline (synthetic batches): ``DetectionModel(dict)`` (``nn/tasks.py:321``) whose ``forward(batch)``
runs ``v8DetectionLoss`` (``utils/loss.py:152``), ``loss.sum()``, GradScaler backward, ``clip_grad_norm_(10)``, the SGD the
trainer's ``build_optimizer`` builds (three groups, nesterov; ``trainer.py:788-830``), ``ModelEMA.update``
(``utils/torch_utils.py:620-672``).  Used by ``bench.py --impl reference`` (CPU, fp32), by its ``gpu_eager_baseline`` leg
(the reference's blocks as they run on a GPU today: PyTorch eager, bf16 autocast or the trainer's native fp16 +
GradScaler) and by tests/test_gpu_plugin.py.  Nothing in the product package imports this file.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import ref_loader

SWIN_DIM = {"n": 128, "s": 256, "m": 384, "l": 512, "x": 640}   # yolov8.yaml:664 author note


def available() -> bool:
    return ref_loader.available()


def build_model(scale: str = "n", nc: int = 80, verbose: bool = False):
    """The fork's active architecture (cfg/models/v8/yolov8.yaml:734-776) at ``scale`` as the reference's DetectionModel."""
    ref_loader.import_ultralytics()
    from ultralytics.cfg import get_cfg
    from ultralytics.nn.tasks import DetectionModel, yaml_model_load

    d = yaml_model_load(os.path.join(ref_loader.REF_ROOT, f"ultralytics/cfg/models/v8/yolov8{scale}.yaml"))
    d["backbone"][7][3] = [SWIN_DIM[scale]]
    d["head"][3][3] = [SWIN_DIM[scale]]
    model = DetectionModel(d, ch=3, nc=nc, verbose=verbose)
    model.args = get_cfg()          # box 7.5 / cls 0.5 / dfl 1.5 (cfg/default.yaml:98-100)
    for k, p in model.named_parameters():
        if ".dfl" in k:
            p.requires_grad_(False)  # trainer.py:243-262
    return model


def build_optimizer(model, lr=0.01, momentum=0.937, decay=5e-4):
    """trainer.py:818-830 with name='SGD': g2 = biases, g1 = norm weights (no decay), g0 = other weights (decay)."""
    g = [], [], []
    bn = tuple(v for k, v in nn.__dict__.items() if "Norm" in k)
    for module_name, module in model.named_modules():
        for param_name, param in module.named_parameters(recurse=False):
            fullname = f"{module_name}.{param_name}" if module_name else param_name
            if "bias" in fullname:
                g[2].append(param)
            elif isinstance(module, bn) or "logit_scale" in fullname:
                g[1].append(param)
            else:
                g[0].append(param)
    opt = torch.optim.SGD(g[2], lr=lr, momentum=momentum, nesterov=True)
    opt.add_param_group({"params": g[0], "weight_decay": decay})
    opt.add_param_group({"params": g[1], "weight_decay": 0.0})
    return opt


class RefTrainer:
    """amp: None (fp32), 'bf16' (autocast, no scaler) or 'fp16' (autocast + GradScaler: the reference trainer's own AMP)."""

    def __init__(self, scale="n", nc=80, device="cpu", amp=None, seed=0, ema=True, channels_last=False):
        ref_loader.import_ultralytics()
        from ultralytics.utils.torch_utils import ModelEMA

        self.device = torch.device(device)
        self.amp = amp
        torch.manual_seed(seed)
        self.model = build_model(scale, nc).to(self.device).train()
        if channels_last:
            self.model = self.model.to(memory_format=torch.channels_last)
        self.opt = build_optimizer(self.model)
        self.ema = ModelEMA(self.model) if ema else None
        self.scaler = torch.amp.GradScaler(self.device.type, enabled=(amp == "fp16"))

    def to_device(self, host_batch):
        return {k: v.to(self.device, non_blocking=True) for k, v in host_batch.items()}

    def step(self, batch):
        """trainer.py:382-399 + optimizer_step (:614-622).  ``batch``: uint8 img + labels on ``self.device``."""
        dt = {"bf16": torch.bfloat16, "fp16": torch.float16}.get(self.amp, torch.bfloat16)
        with torch.autocast(self.device.type, dtype=dt, enabled=self.amp is not None):
            b = dict(batch)
            b["img"] = b["img"].float() / 255                     # detect/train.py:100
            loss, items = self.model(b)                             # BaseModel.forward(dict) -> loss()
            loss = loss.sum()
        self.scaler.scale(loss).backward()
        self.scaler.unscale_(self.opt)
        torch.nn.utils.clip_grad_norm_(self.model.parameters(), max_norm=10.0)
        self.scaler.step(self.opt)
        self.scaler.update()
        self.opt.zero_grad()
        if self.ema:
            self.ema.update(self.model)
        return items
