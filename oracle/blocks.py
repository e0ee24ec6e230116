"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference's CBAM / SwinBlock / SPPF maths.

Nothing in the product package may import this module: it is the *checker* for the CUDA path
(tests/, ``__graft_entry__.smoke()``) and the thing ``bench.py`` times on the host cores for its
``cpu_baseline`` / ``--impl reference`` legs.  It is never the thing shipped.

Parity pin: the reference's own tests hold NO golden vectors for these blocks, so the
oracle is pinned against outputs of the reference itself, imported in the dev container:
``oracle/make_golden.py`` runs the real ``ultralytics.nn.modules.{cbam,swin_block,block}`` classes on
seeded inputs and commits inputs/weights/outputs/gradients under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function below against those fixtures.

Everything is written from primitives (matmul / exp / erf / explicit window scans) rather than by calling
``nn.MultiheadAttention`` / ``nn.LayerNorm`` so that it is an independent statement of the algorithm:

* CBAM          <- ultralytics/nn/modules/cbam.py:5-71
* SwinBlock     <- ultralytics/nn/modules/swin_block.py:8-58 (+ torch ``F.multi_head_attention_forward``)
* SPPF pooling  <- ultralytics/nn/modules/block.py:201-226 (+ torch ``max_pool2d_with_indices`` tie rule)

All functions are differentiable torch code (float32 or float64), so reference *gradients* are obtained
with autograd over the restatement.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------------
# CBAM  (cbam.py:29-38 channel attention, :48-53 spatial attention, :62-71 composition)
# --------------------------------------------------------------------------------------------------
def cbam_channel_attention(x, w1, w2):
    """sigmoid(W2 relu(W1 avg) + W2 relu(W1 max)) -> [B, C] (cbam.py:35-38).

    w1: [r, C] (shared_MLP.0.weight squeezed), w2: [C, r] (shared_MLP.2.weight squeezed); no biases
    (cbam.py:24-26).
    """
    B, C, H, W = x.shape
    flat = x.reshape(B, C, H * W)
    avg = flat.mean(dim=2)  # AdaptiveAvgPool2d(1), cbam.py:8
    mx = flat.max(dim=2).values  # AdaptiveMaxPool2d(1), cbam.py:9
    h_avg = torch.relu(avg @ w1.t())
    h_max = torch.relu(mx @ w1.t())
    return torch.sigmoid(h_avg @ w2.t() + h_max @ w2.t())


def cbam_spatial_attention(x1, wsa):
    """sigmoid(conv_kxk(cat[mean_c, max_c])) -> [B, H, W] (cbam.py:48-53). wsa: [1, 2, k, k], pad k//2."""
    k = wsa.shape[-1]
    s = torch.stack([x1.mean(dim=1), x1.max(dim=1).values], dim=1)  # avg first, cbam.py:49-51
    z = F.conv2d(s, wsa, padding=k // 2)
    return torch.sigmoid(z[:, 0])


def cbam_forward(x, w1, w2, wsa):
    """out = (x*ca)*sa (cbam.py:62-71)."""
    ca = cbam_channel_attention(x, w1, w2)
    x1 = x * ca[:, :, None, None]
    sa = cbam_spatial_attention(x1, wsa)
    return x1 * sa[:, None, :, :]


# --------------------------------------------------------------------------------------------------
# Conv epilogue: BatchNorm2d (+ SiLU) on a conv output (conv.py:65-79; used by SPPF cv1/cv2, block.py:218-219)
# --------------------------------------------------------------------------------------------------
def bn_act_forward(x, gamma, beta, running_mean, running_var, training=True, momentum=0.03, eps=1e-3, silu=True):
    """act(BatchNorm2d(x)) from primitives.  Training: per-channel batch statistics over (B,H,W), biased variance for
    the normalisation, running stats updated with the UNBIASED variance (torch.nn.BatchNorm2d semantics).
    Returns (z, new_running_mean, new_running_var)."""
    B, C, H, W = x.shape
    if training:
        n = B * H * W
        mean = x.mean(dim=(0, 2, 3))
        var = ((x - mean[None, :, None, None]) ** 2).mean(dim=(0, 2, 3))
        rm = (1 - momentum) * running_mean + momentum * mean.detach()
        rv = (1 - momentum) * running_var + momentum * var.detach() * (n / max(n - 1, 1))
    else:
        mean, var, rm, rv = running_mean, running_var, running_mean, running_var
    y = (x - mean[None, :, None, None]) / torch.sqrt(var[None, :, None, None] + eps)
    y = y * gamma[None, :, None, None] + beta[None, :, None, None]
    z = y * torch.sigmoid(y) if silu else y
    return z, rm, rv


# --------------------------------------------------------------------------------------------------
# SwinBlock (swin_block.py:37-58)
# --------------------------------------------------------------------------------------------------
def layer_norm(t, gamma, beta, eps=1e-5):
    mu = t.mean(dim=-1, keepdim=True)
    var = ((t - mu) ** 2).mean(dim=-1, keepdim=True)  # biased variance
    return (t - mu) / torch.sqrt(var + eps) * gamma + beta


def gelu_erf(a):
    return 0.5 * a * (1.0 + torch.erf(a / math.sqrt(2.0)))


def window_partition(x_nhwc, ws):
    """[B, Hp, Wp, C] -> [B*nW, ws*ws, C]; window order (b, wh, ww), token order (row, col). swin_block.py:8-13."""
    B, Hp, Wp, C = x_nhwc.shape
    t = x_nhwc.reshape(B, Hp // ws, ws, Wp // ws, ws, C)
    return t.permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws, C)


def window_reverse(win, ws, Hp, Wp):
    """Inverse of window_partition. swin_block.py:15-20."""
    C = win.shape[-1]
    B = win.shape[0] // ((Hp // ws) * (Wp // ws))
    t = win.reshape(B, Hp // ws, Wp // ws, ws, ws, C)
    return t.permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, C)


def shift_attention_mask(Hp, Wp, ws, shift, dtype, device=None):
    """Additive mask [nW, L, L] of shifted-window attention (the standard Swin construction): label the padded map by
    the 3x3 regions the cyclic shift creates, partition the labels like the tokens, and put -100 between tokens of
    different regions.  EXTENSION: the reference block has no shift; this is our own specification."""
    img = torch.zeros(1, Hp, Wp, 1, dtype=dtype, device=device)
    cnt = 0
    for hs in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for wsl in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img[:, hs, wsl, :] = cnt
            cnt += 1
    lab = window_partition(img, ws).squeeze(-1)  # [nW, L]
    diff = lab[:, None, :] - lab[:, :, None]
    return torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))


def swin_forward(x, p, num_heads=2, ws=7, shift=0):
    """SwinBlock.forward (swin_block.py:37-58) restated.  ``shift`` > 0 (EXTENSION, not in the reference): cyclic shift of
    the padded map by (-shift, -shift) before partitioning, region mask on the scores, shift back before the crop.

    p: dict with the reference state_dict keys: norm1.weight/bias, attn.in_proj_weight [3C,C],
    attn.in_proj_bias [3C], attn.out_proj.weight [C,C], attn.out_proj.bias, norm2.weight/bias,
    mlp.0.weight [4C,C], mlp.0.bias, mlp.2.weight [C,4C], mlp.2.bias.

    Quirks reproduced on purpose: no shift / no relative-position bias / no mask; zero padding
    happens BEFORE LayerNorm so padded tokens become norm1.bias and act as keys/values; the residual is taken
    from the *normalised* tokens: y1 = LN1(t) + MHA(LN1(t)).
    """
    B, C, H, W = x.shape
    pad_h = (ws - H % ws) % ws
    pad_w = (ws - W % ws) % ws
    xp = F.pad(x, (0, pad_w, 0, pad_h))
    Hp, Wp = H + pad_h, W + pad_w
    if shift:
        xp = torch.roll(xp, shifts=(-shift, -shift), dims=(2, 3))
    t = window_partition(xp.permute(0, 2, 3, 1), ws)  # [nWB, L, C]
    n1 = layer_norm(t, p["norm1.weight"], p["norm1.bias"])
    nWB, L, _ = n1.shape
    hd = C // num_heads
    qkv = n1 @ p["attn.in_proj_weight"].t() + p["attn.in_proj_bias"]
    q, k, v = qkv.split(C, dim=-1)

    def heads(z):  # [nWB, L, C] -> [nWB, nh, L, hd]; head h = channels [h*hd, (h+1)*hd)
        return z.reshape(nWB, L, num_heads, hd).permute(0, 2, 1, 3)

    q, k, v = heads(q) * (1.0 / math.sqrt(hd)), heads(k), heads(v)
    s = q @ k.transpose(-1, -2)
    if shift:
        m = shift_attention_mask(Hp, Wp, ws, shift, s.dtype, s.device)  # [nW, L, L], windows ordered (wh, ww)
        s = (s.reshape(B, -1, num_heads, L, L) + m[None, :, None]).reshape(nWB, num_heads, L, L)
    s = s - s.max(dim=-1, keepdim=True).values
    e = torch.exp(s)
    pr = e / e.sum(dim=-1, keepdim=True)
    o = (pr @ v).permute(0, 2, 1, 3).reshape(nWB, L, C)
    a = o @ p["attn.out_proj.weight"].t() + p["attn.out_proj.bias"]
    y1 = n1 + a  # swin_block.py:50-52
    u = layer_norm(y1, p["norm2.weight"], p["norm2.bias"])
    hmid = gelu_erf(u @ p["mlp.0.weight"].t() + p["mlp.0.bias"])
    y2 = y1 + hmid @ p["mlp.2.weight"].t() + p["mlp.2.bias"]
    out = window_reverse(y2, ws, Hp, Wp).permute(0, 3, 1, 2)
    if shift:
        out = torch.roll(out, shifts=(shift, shift), dims=(2, 3))
    return out[:, :, :H, :W]


# --------------------------------------------------------------------------------------------------
# SPPF pooling cascade (block.py:220-226): y_{i+1} = maxpool(y_i, k, stride 1, pad k//2), concat 4 slices
# --------------------------------------------------------------------------------------------------
def maxpool_scan_np(y: np.ndarray, k: int):
    """One stride-1 'same' max-pool on [N, H, W] with torch's exact rule, as explicit loops over the window.

    Rule (ATen max_pool2d_with_indices): implicit -inf padding, window scanned row-major, a
    candidate replaces the running max iff ``val > max`` OR val is NaN; index = flat h*W+w of the winner.
    Returns (values, int32 indices).  Pure numpy, vectorised over N and positions; O(k*k) passes.
    """
    N, H, W = y.shape
    r = k // 2
    best = np.full((N, H, W), -np.inf, dtype=y.dtype)
    idx = np.full((N, H, W), -1, dtype=np.int32)
    hh, ww = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    first = np.ones((N, H, W), dtype=bool)
    for dy in range(-r, r + 1):
        for dx in range(-r, r + 1):
            sh, sw = hh + dy, ww + dx
            valid = (sh >= 0) & (sh < H) & (sw >= 0) & (sw < W)
            shc, swc = np.clip(sh, 0, H - 1), np.clip(sw, 0, W - 1)
            cand = y[:, shc, swc]
            with np.errstate(invalid="ignore"):
                take = valid[None] & ((cand > best) | np.isnan(cand) | first)
            # ATen initialises maxval=-inf and maxindex to the first in-bounds element of the window
            best = np.where(take, cand, best)
            idx = np.where(take, (shc * W + swc)[None].astype(np.int32), idx)
            first = first & ~valid[None]
    return best, idx


def sppf_pool_cascade_np(y0: np.ndarray, k: int):
    """[B, C, H, W] -> (concat [B, 4C, H, W], idx [3, B, C, H, W]) via three explicit window scans."""
    B, C, H, W = y0.shape
    cur = y0.reshape(B * C, H, W)
    outs, idxs = [cur], []
    for _ in range(3):
        cur, ii = maxpool_scan_np(cur, k)
        outs.append(cur)
        idxs.append(ii)
    cat = np.concatenate([o.reshape(B, C, H, W) for o in outs], axis=1)
    return cat, np.stack([i.reshape(B, C, H, W) for i in idxs])


def sppf_pool_cascade(y0: torch.Tensor, k: int):
    """Differentiable torch form of the same cascade (used for gradient oracles and large sizes)."""
    ys = [y0]
    idx = []
    for _ in range(3):
        v, i = F.max_pool2d(ys[-1], k, 1, k // 2, return_indices=True)
        ys.append(v)
        idx.append(i)
    return torch.cat(ys, 1), torch.stack(idx)


def sppf_pool_backward_np(gcat: np.ndarray, idx: np.ndarray):
    """Gradient of the cascade w.r.t. y0 given per-stage argmax maps.

    G3=g3; G2=g2+S3(G3); G1=g1+S2(G2); dy0=g0+S1(G1), S_k(G)[p] = sum_{q: I_k(q)=p} G[q].  float64 accumulate.
    """
    _, B, C, H, W = idx.shape
    g = gcat.reshape(B, 4, C, H * W).astype(np.float64)
    acc = g[:, 3].copy()
    for st in (2, 1, 0):
        ii = idx[st].reshape(B, C, H * W).astype(np.int64)
        nxt = g[:, st].copy()
        flat_n = nxt.reshape(B * C, H * W)
        flat_a = acc.reshape(B * C, H * W)
        flat_i = ii.reshape(B * C, H * W)
        rows = np.repeat(np.arange(B * C), H * W)
        np.add.at(flat_n, (rows, flat_i.ravel()), flat_a.ravel())
        acc = nxt
    return acc.reshape(B, C, H, W)
