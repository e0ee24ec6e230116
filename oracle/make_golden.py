"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz from the REAL reference modules.

Run in the dev container (where /root/reference exists):  ``python -m oracle.make_golden``.
The reference's own tests hold no vectors for CBAM / SwinBlock / SPPF, so these fixtures --
outputs and autograd gradients of the unmodified reference classes on seeded inputs -- are the parity pin
for ``oracle/blocks.py`` and, on the GPU box (where the reference is absent), for the CUDA kernels.
Each file is a flat npz: ``x``, ``gy`` (upstream grad), ``y``, ``gx``, ``w.<state_dict key>``, ``gw.<key>``.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _run(mod, x, seed):
    g = torch.Generator().manual_seed(seed)
    x = x.clone().requires_grad_(True)
    y = mod(x)
    gy = torch.randn(y.shape, generator=g)
    y.backward(gy)
    d = {"x": x.detach().numpy(), "gy": gy.numpy(), "y": y.detach().numpy(), "gx": x.grad.numpy()}
    for k, v in mod.state_dict().items():
        d["w." + k] = v.numpy()
    for k, p in mod.named_parameters():
        if p.grad is not None:
            d["gw." + k] = p.grad.numpy()
    return d


def main():
    os.makedirs(OUT, exist_ok=True)
    ref_loader.import_ultralytics()
    from ultralytics.nn.modules.block import SPPF
    from ultralytics.nn.modules.cbam import CBAM, ChannelAttention, SpatialAttention
    from ultralytics.nn.modules.swin_block import SwinBlock

    # ---- CBAM: lazy (channels=None -> ratio 16), explicit small (ratio 8), odd spatial sizes
    for name, ctor, shape, seed in [
        ("cbam_lazy_c32", lambda: CBAM(), (2, 32, 9, 10), 11),
        ("cbam_c64_r8", lambda: CBAM(64), (2, 64, 5, 7), 12),
        ("cbam_lazy_c256_p5", lambda: CBAM(), (1, 256, 20, 20), 13),
    ]:
        torch.manual_seed(seed)
        m = ctor()
        x = torch.randn(shape)
        with torch.no_grad():
            m(x)  # lazy MLP creation happens inside forward (cbam.py:31-33)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **_run(m, x, seed + 100))
    # sub-modules return the maps only (cbam.py:38,53)
    torch.manual_seed(21)
    ca, sa = ChannelAttention(32, 16), SpatialAttention(3)
    x = torch.randn(2, 32, 6, 5)
    np.savez_compressed(os.path.join(OUT, "cbam_ca_c32.npz"), **_run(ca, x, 121))
    np.savez_compressed(os.path.join(OUT, "cbam_sa_k3.npz"), **_run(sa, x, 122))

    # ---- SwinBlock: padded (9x10 -> 14x14), exact multiple (14x7), ws=8/heads=4, single window smaller than ws
    for name, args, shape, seed in [
        ("swin_c16_pad", (16, 2, 7), (2, 16, 9, 10), 31),
        ("swin_c32_exact", (32, 2, 7), (1, 32, 14, 7), 32),
        ("swin_c32_ws8_h4", (32, 4, 8), (2, 32, 10, 17), 33),
        ("swin_c16_tiny", (16, 2, 7), (1, 16, 3, 4), 34),
        ("swin_c128_p4", (128, 2, 7), (1, 128, 20, 20), 35),
    ]:
        torch.manual_seed(seed)
        m = SwinBlock(*args)
        # non-trivial LN affine + biases so the padded-token (= norm1.bias) path is exercised
        with torch.no_grad():
            for k, p in m.named_parameters():
                if "norm" in k or "bias" in k:
                    p.add_(0.3 * torch.randn_like(p))
        d = _run(m, torch.randn(shape), seed + 100)
        d["args"] = np.array(args)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)

    # ---- SPPF pool cascade with torch's indices (values + argmax are the bit-exact contract)
    for name, k, shape, mode, seed in [
        ("sppf_k5_rand", 5, (2, 8, 9, 11), "rand", 41),
        ("sppf_k7_rand", 7, (1, 8, 20, 20), "rand", 42),
        ("sppf_k5_ties", 5, (2, 4, 12, 7), "ties", 43),
        ("sppf_k7_const_nan", 7, (1, 4, 8, 8), "const_nan", 44),
        ("sppf_k5_small", 5, (1, 3, 2, 3), "rand", 45),
    ]:
        torch.manual_seed(seed)
        m = SPPF(shape[1] * 2, shape[1] * 2, k)
        y0 = torch.randn(shape)
        if mode == "ties":
            y0 = torch.relu(y0).bfloat16().float().round()  # heavy plateaus -> tie rule decides routing
        elif mode == "const_nan":
            y0 = torch.full(shape, 1.5)
            y0[0, 1, 3, 4] = float("nan")
            y0[0, 2] = float("-inf")
            y0[0, 3, 2, 2] = float("inf")
        ys, idx = [y0], []
        for _ in range(3):
            v, i = torch.nn.functional.max_pool2d(ys[-1], m.m.kernel_size, m.m.stride, m.m.padding,
                                                  return_indices=True)
            ys.append(v)
            idx.append(i)
        cat = torch.cat(ys, 1)
        d = {"y0": y0.numpy(), "cat": cat.numpy(), "idx": torch.stack(idx).numpy().astype(np.int32), "k": np.array(k)}
        if mode in ("rand", "ties"):
            g = torch.Generator().manual_seed(seed + 100)
            y0g = y0.clone().requires_grad_(True)
            yy = [y0g]
            yy.extend(m.m(yy[-1]) for _ in range(3))  # block.py:224-225
            gcat = torch.randn(cat.shape, generator=g)
            torch.cat(yy, 1).backward(gcat)
            d["gcat"], d["gy0"] = gcat.numpy(), y0g.grad.numpy()
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)

    # ---- full SPPF module (cv1 -> pool cascade -> cv2), train-mode BN, k=5
    torch.manual_seed(51)
    m = SPPF(16, 16, 5).train()
    d = _run(m, torch.randn(2, 16, 6, 6), 151)
    np.savez_compressed(os.path.join(OUT, "sppf_module_k5.npz"), **d)
    # ---- the reference Conv block (conv -> BatchNorm2d -> SiLU, conv.py:37-79), train mode: pins the fused epilogue
    from ultralytics.nn.modules.conv import Conv

    for name, args, shape, seed in [("conv_k1_c16", (8, 16, 1, 1), (3, 8, 6, 5), 61), ("conv_k3_c32", (16, 32, 3, 2), (2, 16, 9, 9), 62)]:
        torch.manual_seed(seed)
        m = Conv(*args).train()
        with torch.no_grad():
            m.bn.weight.add_(0.3 * torch.randn_like(m.bn.weight))
            m.bn.bias.add_(0.3 * torch.randn_like(m.bn.bias))
        before = {k: v.clone() for k, v in m.state_dict().items()}
        x = torch.randn(shape)
        d = _run(m, x, seed + 100)
        for k, v in before.items():
            d["w0." + k] = v.numpy()       # state BEFORE the forward (running stats are updated by it)
        with torch.no_grad():
            d["conv_out"] = m.conv(x).numpy()
        d["bn_eps"], d["bn_momentum"] = np.array(m.bn.eps), np.array(m.bn.momentum)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    tot = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print("wrote", len(os.listdir(OUT)), "fixtures,", tot // 1024, "KiB")


if __name__ == "__main__":
    main()
