"""Synthetic COCO-shaped batches: seeded uint8 images + 8 boxes per image."""
from __future__ import annotations

import torch

BOXES_PER_IMAGE = 8


def make_batch(batch_size: int, imgsz: int = 640, nc: int = 80, seed: int = 1234, pin: bool = False):
    """Host-side batch: img uint8 U[0,255] [B,3,S,S]; labels: batch_idx sorted, cls U{0..nc-1}, xywh-normalised
    boxes with centre U[0.2,0.8], size U[0.05,0.35] (the recipe SURVEY probe6 found gives finite loss + grads)."""
    g = torch.Generator().manual_seed(seed)
    n = batch_size * BOXES_PER_IMAGE
    img = torch.randint(0, 256, (batch_size, 3, imgsz, imgsz), dtype=torch.uint8, generator=g)
    batch_idx = torch.arange(batch_size).repeat_interleave(BOXES_PER_IMAGE).float()
    cls = torch.randint(0, nc, (n, 1), generator=g).float()
    cxy = 0.2 + 0.6 * torch.rand(n, 2, generator=g)
    wh = 0.05 + 0.30 * torch.rand(n, 2, generator=g)
    batch = {"img": img, "batch_idx": batch_idx, "cls": cls, "bboxes": torch.cat((cxy, wh), 1)}
    if pin:
        batch = {k: v.pin_memory() for k, v in batch.items()}
    return batch


def batch_nbytes(batch) -> int:
    return sum(v.numel() * v.element_size() for v in batch.values())
