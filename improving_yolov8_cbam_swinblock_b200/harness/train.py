"""Synthetic-batch data-parallel training step for the custom YOLOv8-CBAM-Swin graph.

Mirrors the reference's hot loop (engine/trainer.py:367-399,614-622): H2D copy + ``img.float()/255`` (detect/train.py:
90-115), autocast forward, v8 detection loss, ``loss.sum() * world_size`` (trainer.py:386-388; DDP averages grads),
backward (DDP bucketed NCCL all-reduce of gradients only -- the path's ONE exchange step), ``clip_grad_norm_(10)``,
SGD(momentum 0.937, nesterov) with the trainer's three parameter groups (trainer.py:788-830) and the EMA update
(torch_utils.py:657-672).  One process per GPU; nothing but gradients crosses GPUs.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import graph, loss as hloss


def param_groups(model: nn.Module, weight_decay: float = 5e-4):
    """trainer.py:818-827: biases -> g2 (no decay), norm weights -> g1 (no decay), other weights -> g0 (decay)."""
    bn = tuple(v for k, v in nn.__dict__.items() if "Norm" in k)
    g0, g1, g2 = [], [], []
    for mod_name, mod in model.named_modules():
        for pn, p in mod.named_parameters(recurse=False):
            if not p.requires_grad:
                continue
            full = f"{mod_name}.{pn}" if mod_name else pn
            if "bias" in full:
                g2.append(p)
            elif isinstance(mod, bn):
                g1.append(p)
            else:
                g0.append(p)
    return [{"params": g0, "weight_decay": weight_decay}, {"params": g1, "weight_decay": 0.0},
            {"params": g2, "weight_decay": 0.0}]


class EMA:
    """ModelEMA (torch_utils.py:620-672) with foreach updates instead of the per-tensor Python loop."""

    def __init__(self, model, decay=0.9999, tau=2000):
        self.shadow = [p.detach().clone().float() for p in model.state_dict().values() if p.dtype.is_floating_point]
        self.decay, self.tau, self.updates = decay, tau, 0

    @torch.no_grad()
    def update(self, model):
        self.updates += 1
        d = self.decay * (1 - math.exp(-self.updates / self.tau))
        cur = [p.detach() for p in model.state_dict().values() if p.dtype.is_floating_point]
        torch._foreach_mul_(self.shadow, d)
        torch._foreach_add_(self.shadow, cur, alpha=1 - d)


class Trainer:
    def __init__(self, blocks: dict, scale="n", nc=80, device="cuda", amp_dtype=torch.bfloat16, world_size=1,
                 local_rank=0, channels_last=True, lr=0.01, momentum=0.937, seed=0, ema=True):
        self.device = torch.device(device)
        self.world_size = world_size
        self.amp_dtype = amp_dtype
        self.channels_last = channels_last and self.device.type == "cuda"
        torch.manual_seed(seed)  # identical init on every rank
        model = graph.DetectionGraph(blocks, scale, nc)
        for n_, p in model.named_parameters():
            if ".dfl" in n_:
                p.requires_grad_(False)  # trainer.py:243-262 always freezes the DFL conv
        model = model.to(self.device).train()
        if self.channels_last:
            model = model.to(memory_format=torch.channels_last)
        self.raw = model
        self.model = model
        if world_size > 1:
            kw = dict(device_ids=[local_rank]) if self.device.type == "cuda" else {}
            self.model = nn.parallel.DistributedDataParallel(model, gradient_as_bucket_view=True, static_graph=True, **kw)
        self.criterion = hloss.DetectionLoss(nc, model.stride)
        fused = self.device.type == "cuda"
        self.opt = torch.optim.SGD(param_groups(model), lr=lr, momentum=momentum, nesterov=True, fused=fused)
        self.ema = EMA(model) if ema else None
        self.max_boxes = None
        self._graph, self._graph_error, self._static, self._static_items = None, None, None, None

    def to_device(self, host_batch):
        return {k: v.to(self.device, non_blocking=True) for k, v in host_batch.items()}

    # ---- CUDA-graph replay of the step (host-side step overhead) ------------------------------
    def enable_graph(self, dev_batch, warmup: int = 3) -> bool:
        """Capture forward + loss + backward (+ DDP all-reduce) + clip + SGD of one step into a CUDA graph; shapes are
        static for the synthetic batches.  Falls back to eager launches (returns False) if capture is not possible."""
        if self.device.type != "cuda" or self._graph is not None:
            return self._graph is not None
        if self.world_size > 1:
            # capturing DDP's bucketed NCCL all-reduce deadlocked on the 2-GPU box (round 1): data-parallel runs launch eagerly
            self._graph_error = "graph capture is single-process only; DDP steps are launched eagerly"
            return False
        try:
            self._static = {k: v.clone() for k, v in dev_batch.items()}
            cur = torch.cuda.current_stream(self.device)
            side = torch.cuda.Stream(self.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                for _ in range(max(warmup, 11 if self.world_size > 1 else 3)):
                    self._core(self._static)
            cur.wait_stream(side)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._static_items = self._core(self._static)
            self._graph = g
        except Exception as e:  # noqa: BLE001  capture is an optimisation; eager launches remain correct
            self._graph, self._graph_error = None, f"{type(e).__name__}: {e}"
            torch.cuda.synchronize(self.device)
        return self._graph is not None

    def step(self, dev_batch):
        """One optimizer step on a device-resident batch; returns the detached loss items [box, cls, dfl]."""
        if self._graph is not None:
            for k, v in dev_batch.items():
                self._static[k].copy_(v, non_blocking=True)
            self._graph.replay()
            items = self._static_items.clone()
        else:
            items = self._core(dev_batch)
        if self.ema is not None:
            self.ema.update(self.raw)
        return items

    def _core(self, dev_batch):
        img = dev_batch["img"].float() / 255  # detect/train.py:100
        if self.channels_last:
            img = img.contiguous(memory_format=torch.channels_last)
        use_amp = self.amp_dtype is not None and self.amp_dtype != torch.float32
        with torch.autocast(self.device.type, dtype=self.amp_dtype or torch.bfloat16, enabled=use_amp):
            feats = self.model(img)
            loss, items = self.criterion([f.float() for f in feats], dev_batch, max_boxes=self.max_boxes)
        total = loss.sum()
        if self.world_size > 1:
            total = total * self.world_size
        total.backward()
        torch.nn.utils.clip_grad_norm_(self.raw.parameters(), max_norm=10.0, foreach=True)
        self.opt.step()
        self.opt.zero_grad(set_to_none=True)
        return items

    def step_from_host(self, host_batch):
        """The user-facing call: pinned host batch in, host loss items out (H2D + D2H inside)."""
        items = self.step(self.to_device(host_batch))
        return items.cpu()
