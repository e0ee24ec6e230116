import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: d[k] for k in d.files}


def state_dict_of(g):
    return {k[2:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("w.")}


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a-b|| / ||b|| in float64 (the 'relative error' of the bf16 tolerance in BASELINE.json)."""
    a, b = a.detach().double().flatten().cpu(), b.detach().double().flatten().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def assert_close_f32(a, b, what, rtol=1e-5, atol=2e-6):
    a, b = a.float().cpu(), b.float().cpu()
    scale = float(b.abs().max().clamp_min(1.0))
    bad = (a - b).abs() > (atol * scale + rtol * b.abs())
    assert not bool(bad.any()), f"{what}: {int(bad.sum())} / {bad.numel()} elements off, max abs {float((a - b).abs().max()):.3e}"


def to_cl(x):
    return x.contiguous(memory_format=torch.channels_last)
