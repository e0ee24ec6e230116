"""Per-kernel SASS marker counts of libb200yolo.so (what proves a Blackwell-native kernel, B200_PROFILING.md):
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor load/store, UBLKCP = cp.async.bulk,
HMMA = legacy mma.sync: 0 in every SwinBlock / GEMM kernel (those are tcgen05); the narrow-channel convolution kernels of the
callers (stem 3 -> 16, 3x3 wgrad / stride-2 dgrad at 16 channels: HBM-bound, K = 27..288, tensor work negligible) use it on purpose.  usage: python profiles/sass_markers.py > profiles/sass_markers_rNN.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "improving_yolov8_cbam_swinblock_b200", "libb200yolo.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
MARK = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "MUFU.TANH"]
cur, counts, order = None, collections.defaultdict(collections.Counter), []
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        order.append(cur)
        continue
    if cur is None:
        continue
    for k in MARK:
        if re.search(r"\b" + re.escape(k), line):
            counts[cur][k] += 1
    if re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
        counts[cur]["_instr"] += 1
demangled = subprocess.run(["c++filt"], input="\n".join(order), capture_output=True, text=True).stdout.splitlines()
print("| kernel | SASS instr | " + " | ".join(MARK) + " |")
print("|---|---:|" + "---:|" * len(MARK))
tot = collections.Counter()
rows = []
for name, dem in zip(order, demangled):
    c = counts[name]
    if not any(c[k] for k in MARK if k not in ("SYNCS", "MUFU.TANH")):
        continue
    short = dem.replace("(anonymous namespace)::", "").replace("void ", "")
    short = re.sub(r"\(.*", "", short)
    rows.append((short, c))
    tot.update(c)
for short, c in sorted(rows):
    print(f"| `{short[:90]}` | {c['_instr']} | " + " | ".join(str(c[k]) for k in MARK) + " |")
print(f"| **total (kernels listed)** | {tot['_instr']} | " + " | ".join(str(tot[k]) for k in MARK) + " |")
