"""Where does a warp-specialised kernel wait?  Reads `ncu -i rep --page source --csv` (SASS view) and lists every mbarrier
try-wait site (SYNCS.PHASECHK...TRYWAIT) with its shared-memory offset, the number of times it executed (= spin count) and
the stall samples on / around it; plus the overall instruction and sample totals.
usage: ncu -i x.ncu-rep --page source --csv > x.csv ; python profiles/ncu_waits.py x.csv [bar_base_hex name0 name1 ...]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
hdr = rows[hi]
ia, isrc, isamp, iins = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = [(r[ia], r[isrc].strip(), int(r[isamp]), int(r[iins])) for r in rows[hi + 1:] if len(r) > iins and r[isamp].isdigit()]
base = int(sys.argv[2], 16) if len(sys.argv) > 2 else None
names = sys.argv[3:]
tot_i, tot_s = sum(d[3] for d in data), sum(d[2] for d in data)
print(f"warp instructions {tot_i}, stall samples {tot_s}, static SASS {len(data)}")
agg = {}
for idx, (a, s, smp, ins) in enumerate(data):
    if "TRYWAIT" in s:
        off = s.split("+0x")[-1].split("]")[0]
        near = sum(d[2] for d in data[max(idx - 1, 0):idx + 3])
        e = agg.setdefault(off, [0, 0])
        e[0] += ins
        e[1] += near
for off, (ins, smp) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    label = ""
    if base is not None:
        k = (int(off, 16) - base) // 8
        label = names[k] if 0 <= k < len(names) else f"bar[{k}]"
    print(f"  +0x{off} {label:12s} tries {ins:9d}  samples {smp}")
