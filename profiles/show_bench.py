"""Pretty-print a bench.py JSON line: headline + per-kernel roofline table."""
import json
import sys

d = json.load(open(sys.argv[1]))
print(f"value {d['value']:.1f} {d['unit']}  {d['ms_per_step']:.2f} ms/step   e2e {d['e2e']['value']:.1f}   launches {d.get('gpu_launches')}  clocks {d.get('clocks')}")
r = d.get("roofline") or {}
print({k: r[k] for k in r if k not in ("all_kernels", "kernels_ms_per_step")})
for k, v in sorted((r.get("all_kernels") or {}).items(), key=lambda kv: -kv[1]["avg_ms"] * kv[1]["calls_per_step"]):
    print(f"{v['avg_ms'] * v['calls_per_step']:8.3f} ms/step {v['avg_ms'] * 1e3:8.1f} us x{v['calls_per_step']:4.1f}  frac={v.get('frac')} {v.get('bound')}  {k}")
for m in d.get("modules", []):
    print(f"  module {m['kernel']:24s} {m['ms'] * 1e3:8.1f} us  frac={m['frac']}  {m['bound']} {m['shape']}")
if "cpu_baseline" in d:
    print("cpu_baseline", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], "cores")
