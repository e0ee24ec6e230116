"""Attribute executed SASS instructions / stall samples of one kernel to source-line ranges ("phases").

usage: ncu -i rep --page source --csv --print-source cuda,sass > x.csv ; python profiles/ncu_phase_split.py x.csv file.cu L1:name L2:name ...
Instructions inlined from headers are attributed to the closest preceding (by address) instruction of `file.cu`."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
main = sys.argv[2]
marks = sorted((int(a.split(":")[0]), a.split(":")[1]) for a in sys.argv[3:])
sass = []  # (addr, is_main, line, instr, samples)
cur_file, cur_line = None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1]
        continue
    if r[0].strip().isdigit():
        cur_line = int(r[0])
        continue
    if r[0] == "" and len(r) > 7 and r[2].startswith("0x"):
        try:
            sass.append((int(r[2], 16), cur_file.endswith(main), cur_line, int(r[7]), int(r[6]), r[3].strip()))
        except ValueError:
            pass
sass.sort()
tot_i = sum(s[3] for s in sass)
tot_s = sum(s[4] for s in sass)
buckets = {}
line = 0
for addr, is_main, ln, ins, smp, txt in sass:
    if is_main:
        line = ln
    name = "pre"
    for l, n in marks:
        if line >= l:
            name = n
    b = buckets.setdefault(name, [0, 0, 0])
    b[0] += ins
    b[1] += smp
    b[2] += 1
print(f"total warp-instr {tot_i}  samples {tot_s}  static SASS {len(sass)}")
for name, (i, s, n) in buckets.items():
    print(f"{name:12s} instr {i:9d} {100 * i / tot_i:5.1f}%   samples {100 * s / max(tot_s, 1):5.1f}%   static {n}")
