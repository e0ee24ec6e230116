"""Per-module roofline sweep (BASELINE.json config 5 / SURVEY section 8(d)): each hand-written kernel alone, CUDA
events on the launching stream, L2 flushed (256 MB write) between launches, algorithmic work / measured time."""
from __future__ import annotations

import torch

CHANNELS = {"n": (64, 128, 256), "s": (128, 256, 512), "m": (192, 384, 576)}


def _shapes(scale, B):
    c3, c4, c5 = CHANNELS[scale]
    return {"P3": (B, c3, 80, 80), "P4": (B, c4, 40, 40), "P5": (B, c5, 20, 20)}


def algorithmic_work(scale: str, B: int, s: int = 2, ws: int = 7):
    """ABI entry point -> algorithmic bytes / FLOPs per launch at the MODEL's shapes (SURVEY section 8(d) figures)."""
    _, c4, c5 = CHANNELS[scale]
    n5 = B * c5 * 400
    n0 = B * (c5 // 2) * 400
    tok_real = B * 1600
    nw = -(-40 // ws)
    T = B * nw * nw * ws * ws
    hbm = lambda amount, note: {"bound": "hbm", "amount": float(amount), "note": note}  # noqa: E731
    return {
        "b200_cbam_fwd": hbm(2 * n5 * s, "2*N*s: read x once, write out once (N=B*C5*H*W)"),
        "b200_cbam_bwd": hbm(3 * n5 * s, "3*N*s: read x, read g_out, write g_x"),
        "b200_sppf_pool_fwd": hbm(5 * n0 * s, "5*N0*s: read y0, write the 4*c_ concat (N0=B*c_*H*W)"),
        "b200_sppf_pool_bwd": hbm(6 * n0 * s, "6*N0*s: read 4*c_ grad + y0, write g_y0"),
        "b200_swin_ln1_partition": hbm((tok_real + T) * c4 * s, "read x (real tokens) + write n1 (padded tokens)"),
        "b200_swin_attn_fwd": hbm(4 * T * c4 * s, "standalone attention is HBM-bound: read qkv (3*T*C*s) + write o (T*C*s); "
                                  f"FLOPs {tok_real * 4 * ws * ws * c4 / 1e9:.2f} G on un-padded tokens"),
        "b200_swin_attn_bwd": hbm(8 * T * c4 * s, "read qkv, o, g_o (5*T*C*s) + write g_qkv (3*T*C*s)"),
        "b200_swin_attn_fwd_tc": hbm(4 * T * c4 * s, "stand-alone attention is HBM-bound (22 FLOP/B): read qkv + write o = 4*T*C*s"),
        "b200_swin_attn_bwd_tc": hbm(7 * T * c4 * s, "read qkv, g_o (4*T*C*s) + write g_qkv (3*T*C*s)"),
        "b200_swin_res_ln2": hbm(2 * T * c4 * s, "read y1, write u (the residual add is fused into the out_proj GEMM epilogue)"),
        "b200_swin_gelu": hbm(2 * T * 4 * c4 * s, "read a, write h (backward: +1 read)"),
        "b200_swin_res_reverse": hbm((2 * T + tok_real) * c4 * s, "read y1, m; write out (real tokens)"),
        "b200_swin_partition": hbm((tok_real + T) * c4 * s, "read g (real), write token-major g"),
        "b200_swin_ln_bwd": hbm(4 * T * c4 * s, "read g_out, x, g_res; write g_in"),
        "b200_colsum": hbm(T * c4 * s, "read the activation-gradient matrix once (C..4C columns)"),
    }


def bn_work(tag: str, s: int = 2):
    """Roofline entry for a shape-tagged Conv-epilogue launch: b200_bn_silu_{fwd,bwd}[rows x C]."""
    import re

    m = re.match(r"b200_bn_silu_(fwd|bwd)\[(\d+)x(\d+)\]", tag)
    if not m:
        return None
    n = int(m.group(2)) * int(m.group(3))
    if m.group(1) == "fwd":
        return {"bound": "hbm", "amount": 2.0 * n * s, "note": "2*N*s: read the conv output once, write act(bn(.)) once (the kernel reads it twice: statistics, apply)"}
    return {"bound": "hbm", "amount": 3.0 * n * s, "note": "3*N*s: read x and g once, write g_x once (the kernel reads both twice: reduce, apply)"}


def seam_work(tag: str, s: int = 2):
    """Roofline entry for a shape-tagged channel-concat launch: b200_nhwc_concat[rows x C_total]."""
    import re

    m = re.match(r"b200_nhwc_concat\[(\d+)x(\d+)\]", tag)
    if m:
        n = int(m.group(1)) * int(m.group(2))
        return {"bound": "hbm", "amount": 2.0 * n * s, "note": "2*N*s: every source element read once, written once"}
    m = re.match(r"b200_nhwc_add\[(\d+)x(\d+)x(\d+)\]", tag)
    if m:   # gradient fan-in: n sources read once, the dense sum written once
        n, k = int(m.group(1)) * int(m.group(2)), int(m.group(3))
        return {"bound": "hbm", "amount": (k + 1.0) * n * s, "note": "(n+1)*N*s: n gradient maps read once, their sum written once"}
    # the callers' convolutions served by hand-written kernels: one pass over the operands (the tensor-core work is negligible
    # at 3 / 16 input channels: these launches are bounded by HBM)
    m = re.match(r"b200_conv3x3_dgrad_s2\[(\d+)x(\d+)x(\d+)x(\d+)<-(\d+)\]", tag)
    if m:
        B, H, W, cin, cout = (int(v) for v in m.groups())
        return {"bound": "hbm", "amount": (B * H * W * cin + B * (H // 2) * (W // 2) * cout) * float(s),
                "note": "read gy once, write gx once (weights negligible)"}
    m = re.match(r"b200_conv3x3_fwd_s2\[(\d+)x(\d+)x(\d+)x(\d+)->(\d+)\]", tag)
    if m:
        B, H, W, cin, cout = (int(v) for v in m.groups())
        return {"bound": "hbm", "amount": (B * H * W * cin + B * (H // 2) * (W // 2) * cout) * float(s),
                "note": "read x once, write y once (weights negligible)"}
    m = re.match(r"b200_conv3x3_wgrad\[(\d+)x(\d+)x(\d+)x(\d+)->(\d+),s(\d)\]", tag)
    if m:
        B, H, W, cin, cout, st = (int(v) for v in m.groups())
        return {"bound": "hbm", "amount": (B * H * W * cin + B * (H // st) * (W // st) * cout) * float(s),
                "note": "read x and gy once (the weight gradient is a few KB)"}
    m = re.match(r"b200_stem_conv_(fwd|wgrad)\[(\d+)x(\d+)x(\d+)x(\d+)->(\d+)\]", tag)
    if m:
        B, H, W, cin, cout = (int(v) for v in m.groups()[1:])
        return {"bound": "hbm", "amount": (B * H * W * cin + B * (H // 2) * (W // 2) * cout) * float(s),
                "note": "one pass over the input and the output (forward) / the output gradient (weight gradient)"}
    return None


def fused_work(tag: str):
    """Roofline entry for the fused SwinBlock launches: tensor-core FLOPs on UN-PADDED tokens (the SwinBlock is
    bounded by the tensor pipe; padding / recompute show up as lost efficiency)."""
    import re

    m = re.match(r"b200_swin_mlp_(fwd|bwd)\[(\d+)x(\d+)\]", tag)
    if m:
        rows, C = int(m.group(2)), int(m.group(3))
        if m.group(1) == "fwd":
            return {"bound": "tensor", "amount": 16.0 * rows * C * C, "note": "mlp.0 + mlp.2: 2 x 2*rows*C*4C FLOP (LN2 / GELU / residual ride along)"}
        return {"bound": "tensor", "amount": 24.0 * rows * C * C,
                "note": "recompute mlp.0, d hidden = g W2, d xhat = g_a W1: 3 x 2*rows*C*4C FLOP (the two weight gradients are b200_gemm_splitk launches)"}
    m = re.match(r"b200_bce_logits_(fwd|bwd)\[(\d+)x(\d+)\]", tag)
    if m:
        n = int(m.group(2)) * int(m.group(3))
        if m.group(1) == "fwd":
            return {"bound": "hbm", "amount": 2.0 * n, "note": "N*s: the class logits read once (16-bit)"}
        return {"bound": "hbm", "amount": 4.0 * n, "note": "2*N*s: logits read once, gradient written once"}
    m = re.match(r"b200_swin_attn_block_fwd\[(\d+)x(\d+)x(\d+)x(\d+),ws(\d+)\]", tag)
    if m:
        B, C, H, W, ws = (int(v) for v in m.groups())
        return {"bound": "tensor", "amount": float(B * H * W) * (8.0 * C * C + 4.0 * ws * ws * C),
                "note": "in_proj + QK^T + PV + out_proj on un-padded tokens: B*H*W*(8C^2 + 4*ws^2*C) FLOP"}
    return None


def ncu_kernel_name(tag: str):
    """ABI entry point (or shape-tagged GEMM) -> kernel instantiation name used in profiles/traffic_rNN.json."""
    import re

    fixed = {"b200_cbam_fwd": "b200::cbam_cluster_fwd_kernel<__nv_bfloat16, 8, 1>",
             "b200_sppf_pool_fwd": "b200::sppf_pool_fwd_kernel<__nv_bfloat16, 5, 16, 0>",
             "b200_sppf_pool_bwd": "b200::sppf_pool_bwd_inplace_kernel<__nv_bfloat16, 5, 8, 320>",
             "b200_swin_attn_fwd_tc": "b200::swin_attn_fwd_tc_kernel<64>", "b200_swin_attn_bwd_tc": "b200::swin_attn_bwd_tc_kernel<64>",
             "b200_swin_ln1_partition": "b200::swin_ln1_partition_vec_kernel<__nv_bfloat16, 8, 2>",
             "b200_swin_res_ln2": "b200::swin_res_ln2_vec_kernel<__nv_bfloat16, 8, 2>",
             "b200_swin_ln_bwd": "b200::swin_ln_bwd_vec_kernel<__nv_bfloat16, 8, 2, 0>",   # the norm2 instance (with the residual gradient)
             "b200_swin_res_reverse": "b200::swin_move_vec_kernel<__nv_bfloat16, 8, 2, 0>",
             "b200_swin_partition": "b200::swin_move_vec_kernel<__nv_bfloat16, 8, 2, 1>"}
    if tag in fixed:
        return fixed[tag]
    for pre, name in (("b200_swin_mlp_fwd[", "b200::swin_mlp_fwd_kernel<1>"), ("b200_swin_mlp_bwd[", "b200::swin_mlp_bwd_kernel<1>"),
                      ("b200_swin_attn_block_fwd[", "b200::swin_attn_block_fwd_kernel<1>")):
        if tag.startswith(pre):
            return name
    m = re.match(r"b200_gemm_nt\[(\d+)x(\d+)x(\d+),epi(\d)\]", tag)
    if m:  # the dispatch table of b200_gemm_nt (csrc/gemm_tc.cu): <BLOCK_N, STAGES, EPI, epilogue warps, staging buffers>
        n, epi = int(m.group(2)), int(m.group(4))
        if n >= 128:
            return f"b200::gemm_nt_kernel<128, {4 if epi == 0 else 3}, {epi}, 8, 2>"
        return f"b200::gemm_nt_kernel<64, 4, {epi}, 8, 1>"
    return None


def gemm_work(tag: str, peaks: dict, s: int = 2):
    """Roofline entry for a tagged GEMM launch: b200_gemm_nt[MxNxK,epiE] / b200_gemm_splitk[MxNxK]."""
    import re

    m = re.match(r"b200_gemm_(nt|splitk)\[(\d+)x(\d+)x(\d+)(?:,epi(\d))?\]", tag)
    if not m:
        return None
    kind, M, N, K, epi = m.group(1), int(m.group(2)), int(m.group(3)), int(m.group(4)), int(m.group(5) or 0)
    flops = 2.0 * M * N * K
    if kind == "nt":
        byts = (M * K + N * K + M * N * (2 if epi in (1, 2, 3) else 1)) * s
    else:
        byts = (M * K + N * K) * s + M * N * 4
    t_hbm, t_tc = byts / (peaks["hbm_gbs"] * 1e9), flops / (peaks["bf16_tflops_sustained"] * 1e12)
    if t_hbm >= t_tc:
        return {"bound": "hbm", "amount": byts, "note": f"operands + outputs once ({byts / 1e6:.1f} MB; {flops / 1e9:.1f} GFLOP)"}
    return {"bound": "tensor", "amount": flops, "note": f"2*M*N*K ({flops / 1e9:.1f} GFLOP; {byts / 1e6:.1f} MB)"}


def _time(fn, iters, flush):
    """median ms of fn's GPU work: L2 flushed, and a ~1 ms spin kernel queued first so that the host-side launch
    latency of fn (Python + ctypes) is not inside the event pair."""
    ms = []
    for _ in range(iters + 3):
        flush.zero_()
        torch.cuda._sleep(2_000_000)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        ms.append(s.elapsed_time(e))
    ms = sorted(ms[3:])
    return ms[len(ms) // 2]


def run(scale: str, B: int, peaks: dict, iters: int = 10):
    """-> list of {kernel, shape, dtype, ms, achieved, peak, frac, bound} for CBAM / SPPF pool / SwinBlock."""
    from .. import functional as Fb
    from .. import modules as M

    dev = torch.device("cuda", torch.cuda.current_device())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sh = _shapes(scale, B)
    out = []
    dt = torch.bfloat16
    hbm, tf = peaks["hbm_gbs"], peaks["bf16_tflops"]

    def rec(kernel, shape, ms, amount, bound, note=""):
        ach = amount / (ms * 1e-3) / (1e9 if bound == "hbm" else 1e12)
        peak = hbm if bound == "hbm" else tf
        out.append({"kernel": kernel, "shape": list(shape), "dtype": "bf16", "ms": round(ms, 5), "bound": bound,
                    "achieved": round(ach, 2), "peak": peak, "unit": "GB/s" if bound == "hbm" else "TFLOP/s",
                    "frac": round(ach / peak, 4), "note": note})

    torch.manual_seed(0)
    # ---- CBAM at P5 (the model's use), P4 and P3 (BASELINE configs[3]/[4]: CBAM at P3/P4/P5)
    for lvl in ("P5", "P4", "P3"):
        shape = sh[lvl]
        x = torch.randn(shape, device=dev).to(dt).contiguous(memory_format=torch.channels_last)
        mod = M.CBAM()
        mod(torch.zeros(1, shape[1], 2, 2))
        mod = mod.to(dev)
        n = x.numel()
        with torch.no_grad():
            rec("cbam_fwd", shape, _time(lambda: mod(x), iters, flush), 2 * n * 2, "hbm", lvl)
        xg = x.clone().requires_grad_(True)
        y = mod(xg)
        g = torch.randn_like(y)
        rec("cbam_bwd", shape, _time(lambda: torch.autograd.grad(y, xg, g, retain_graph=True), iters, flush), 3 * n * 2, "hbm", lvl)
        del y, xg
    # ---- Conv epilogue (BatchNorm2d + SiLU) at the SPPF cv1 shape and at a P3 Conv shape
    import torch.nn as nn

    for lvl, shape in (("P5 SPPF.cv1", (sh["P5"][0], sh["P5"][1] // 2, 20, 20)), ("P3 Conv", sh["P3"])):
        bn = nn.BatchNorm2d(shape[1], eps=1e-3, momentum=0.03).to(dev).train()
        x = torch.randn(shape, device=dev).to(dt).contiguous(memory_format=torch.channels_last)
        n = x.numel()
        with torch.no_grad():
            rec("bn_silu_fwd", shape, _time(lambda: Fb.bn_act(x, bn, True), iters, flush), 2 * n * 2, "hbm", lvl)
        xg = x.clone().requires_grad_(True)
        z = Fb.bn_act(xg, bn, True)
        g = torch.randn_like(z)
        rec("bn_silu_bwd", shape, _time(lambda: torch.autograd.grad(z, xg, g, retain_graph=True), iters, flush), 3 * n * 2, "hbm", lvl)
        del z, xg
    # ---- SPPF pool at P5, k = 5 and 7
    Bc, c5, H, W = sh["P5"]
    y0 = torch.randn((Bc, c5 // 2, H, W), device=dev).to(dt).contiguous(memory_format=torch.channels_last)
    n0 = y0.numel()
    for k in (5, 7):
        with torch.no_grad():
            rec(f"sppf_pool_fwd_k{k}", y0.shape, _time(lambda: Fb.sppf_pool(y0, k), iters, flush), 5 * n0 * 2, "hbm", "P5")
        yg = y0.clone().requires_grad_(True)
        cat = Fb.sppf_pool(yg, k)
        g = torch.randn_like(cat)
        rec(f"sppf_pool_bwd_k{k}", y0.shape, _time(lambda: torch.autograd.grad(cat, yg, g, retain_graph=True), iters, flush),
            6 * n0 * 2, "hbm", "P5")
        del cat, yg
    # ---- seams: channel concat at a P3 C2f, nearest 2x up-sampling P4 -> P3, uint8 -> NHWC input conversion
    Bq, c3 = sh["P3"][0], sh["P3"][1]
    wide = torch.randn((Bq, c3, 80, 80), device=dev).to(dt).contiguous(memory_format=torch.channels_last)
    parts = list(wide.chunk(2, 1)) + [torch.randn((Bq, c3 // 2, 80, 80), device=dev).to(dt).contiguous(memory_format=torch.channels_last)]
    ncat = sum(p_.numel() for p_ in parts)
    with torch.no_grad():
        rec("nhwc_concat", (Bq, 3 * (c3 // 2), 80, 80), _time(lambda: Fb.nhwc_concat(parts), iters, flush), 2 * ncat * 2, "hbm",
            "P3 C2f: two chunk slices + one bottleneck output")
        xu = torch.randn(sh["P4"], device=dev).to(dt).contiguous(memory_format=torch.channels_last)
        rec("nhwc_upsample_fwd", sh["P4"], _time(lambda: Fb.nhwc_upsample_nearest(xu, 2, 2), iters, flush), 5 * xu.numel() * 2, "hbm",
            "P4 -> P3, nearest x2")
        img = torch.randint(0, 256, (Bq, 3, 640, 640), dtype=torch.uint8, device=dev)
        rec("u8_to_nhwc", (Bq, 3, 640, 640), _time(lambda: Fb.u8_to_nhwc(img, dt, 255.0), iters, flush), img.numel() * 3, "hbm",
            "uint8 NCHW -> bf16 NHWC / 255")
    del wide, parts, xu, img
    # ---- SwinBlock at P4 (whole block vs tensor peak; FLOPs counted on un-padded tokens)
    Bs, c4, H, W = sh["P4"]
    for ws, shift in ((7, 0), (7, 3), (8, 0), (8, 4)):   # configs[4]: window 7 and 8, shift off / on (shift = extension)
        blk = M.SwinBlock(c4, 2, ws, shift).to(dev)
        x = torch.randn((Bs, c4, H, W), device=dev).to(dt).contiguous(memory_format=torch.channels_last)
        flops = Bs * H * W * (24 * c4 * c4 + 4 * ws * ws * c4)
        with torch.no_grad(), torch.autocast("cuda", dtype=dt):
            tag = f"ws{ws}" + (f"_shift{shift}" if shift else "")
            rec(f"swin_block_fwd_{tag}", x.shape, _time(lambda: blk(x), iters, flush), flops, "tensor", "P4, whole block")
        xg = x.clone().requires_grad_(True)
        with torch.autocast("cuda", dtype=dt):
            y = blk(xg)
        g = torch.randn_like(y)
        params = [xg] + list(blk.parameters())
        rec(f"swin_block_bwd_{tag}", x.shape, _time(lambda: torch.autograd.grad(y, params, g, retain_graph=True), iters, flush),
            2 * flops, "tensor", "P4, whole block backward (2x fwd FLOPs)")
        del y, xg
    return out
