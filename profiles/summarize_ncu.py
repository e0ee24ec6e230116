"""Turn ncu CSV exports into the small tables committed under profiles/.

  python profiles/summarize_ncu.py launches gpurun_out/launches_step_r01.csv profiles/step_launches_r01.md
  python profiles/summarize_ncu.py kernels  gpurun_out/mods_r01_raw.csv      profiles/ncu_kernels_r01.md profiles/traffic_r01.json
"""
import collections
import csv
import json
import re
import sys


def _num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return float("nan")


MINE = re.compile(r"\b(swin_mlp_\w+_kernel|swin_attn_block_\w+_kernel|cbam_\w+_kernel|fold_partials_kernel|bn_\w+_kernel|sppf_pool_(?:fwd|bwd)(?:_inplace)?_kernel|swin_\w+_kernel|gemm_nt_kernel|"
                  r"gemm_splitk_kernel|fold_splits_kernel|fold_ln_kernel|fold_rows_kernel|colsum_partial_kernel|nhwc_concat_kernel|"
                  r"u8_to_nhwc_kernel|upsample_(?:fwd|bwd)_kernel)\b")


def short(name):
    m = MINE.search(name)
    if m:
        t = re.search(re.escape(m.group(1)) + r"<([^()]*?)>(?:\(|$)", name)
        return "b200::" + m.group(1) + (f"<{t.group(1)}>" if t else "")
    return re.sub(r"\(.*", "", re.sub(r"^void ", "", name))[:80]


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hi]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        v = _num(r[vi])
        u = r[ui]
        v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
        k = short(r[ki])
        k = re.sub(r"<.*", "", k) if not k.startswith("b200::") else k
        tot[k] += v
        cnt[k] += 1
    T = sum(tot.values())
    mine = sum(v for k, v in tot.items() if k.startswith("b200::"))
    with open(dst, "w") as f:
        f.write(f"# One training step (B=64, 640^2, bf16) under `ncu --metrics gpu__time_duration.sum` (cold-cache, serialised)\n\n")
        f.write(f"{sum(cnt.values())} launches, {T / 1e3:.2f} ms of kernel time; hand-written (b200::) kernels: "
                f"{sum(c for k, c in cnt.items() if k.startswith('b200::'))} launches, {mine / 1e3:.2f} ms = {100 * mine / T:.1f} % of the step.\n\n")
        f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, v in tot.most_common(45):
            f.write(f"| `{k}` | {cnt[k]} | {v:.1f} | {100 * v / T:.1f} % |\n")
    print("wrote", dst)


def kernels(src, dst, traffic_dst):
    rows = list(csv.reader(open(src)))
    while rows and "Kernel Name" not in rows[0]:   # `--log-file` output starts with ncu's ==PROF== lines
        rows.pop(0)
    h = rows[0]
    want = {"dur_us": "gpu__time_duration.sum", "dram_rd": "dram__bytes_read.sum", "dram_wr": "dram__bytes_write.sum",
            "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "tensor_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "tensor_inst": "sm__inst_executed_pipe_tensor.sum",
            "occ_pct": "sm__warps_active.avg.pct_of_peak_sustained_active", "regs": "launch__registers_per_thread",
            "grid": "launch__grid_size", "block": "launch__block_size", "cluster": "launch__cluster_size"}
    idx = {k: h.index(v) for k, v in want.items() if v in h}
    units = rows[1]
    ki = h.index("Kernel Name")
    out, traffic = [], {}
    for r in rows[2:]:
        d = {k: _num(r[i]) for k, i in idx.items()}
        for k in ("dram_rd", "dram_wr"):
            if k in idx:
                u = units[idx[k]]
                d[k] *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        if "dur_us" in idx:
            u = units[idx["dur_us"]]
            d["dur_us"] *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}.get(u, 1)
        d["name"] = short(r[ki])
        out.append(d)
        traffic.setdefault(d["name"], []).append(d.get("dram_rd", 0) + d.get("dram_wr", 0))
    with open(dst, "w") as f:
        f.write("# `ncu --set full --clock-control none` of every hand-written kernel at the model's shapes (B=64, bf16)\n\n")
        f.write("Durations are cold-cache, serialised replays; DRAM bytes are per launch.  Inputs of these kernels were just\n"
                "written by the previous kernel, so part of the algorithmic traffic is served by the 126 MB L2 (DRAM bytes\n"
                "below the algorithmic figure are expected).\n\n")
        f.write("| kernel | grid x block (cluster) | regs | us | DRAM rd MB | DRAM wr MB | DRAM % | SM % | tensor pipe % | warps active % |\n")
        f.write("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        for d in out:
            f.write(f"| `{d['name']}` | {int(d.get('grid', 0))} x {int(d.get('block', 0))} ({int(d.get('cluster', 0))}) | {int(d.get('regs', 0))} | "
                    f"{d.get('dur_us', 0):.1f} | {d.get('dram_rd', 0) / 1e6:.1f} | {d.get('dram_wr', 0) / 1e6:.1f} | {d.get('dram_pct', 0):.1f} | "
                    f"{d.get('sm_pct', 0):.1f} | {d.get('tensor_pct', float('nan')):.1f} | {d.get('occ_pct', 0):.1f} |\n")
    # per-launch DRAM bytes per kernel instantiation; null when launches of one instantiation differ (different shapes)
    json.dump({k: (sum(v) / len(v) if max(v) <= 1.1 * max(min(v), 1.0) else None) for k, v in traffic.items()},
              open(traffic_dst, "w"), indent=1)
    print("wrote", dst, traffic_dst)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        kernels(sys.argv[2], sys.argv[3], sys.argv[4])
