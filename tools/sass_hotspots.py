"""Aggregate an `ncu --page source --csv` (SASS view) export: samples per opcode and the hottest instructions.
usage: ncu -i rep.ncu-rep --page source --csv --kernel-name regex:NAME --launch-skip N --launch-count 1 > k.csv
       python tools/sass_hotspots.py k.csv [top_n]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
samp, src = col["# Samples"], col["Source"]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
total = 0
by_op = collections.Counter()
by_op_n = collections.Counter()
by_stall = collections.Counter()
inst = []
exe = col["Instructions Executed"]
for idx, r in enumerate(rows[hi + 1:]):
    if len(r) <= samp or not r[samp]:
        continue
    try:
        n = int(float(r[samp]))
    except ValueError:
        continue
    total += n
    op = r[src].split()[0] if not r[src].startswith("@") else r[src].split()[1]
    op = op.split(".")[0]
    by_op[op] += n
    by_op_n[op] += int(float(r[exe] or 0))
    st = {s: int(float(r[col[s]] or 0)) for s in stalls}
    for s, v in st.items():
        by_stall[s] += v
    inst.append((n, idx, r[src][:90], max(st, key=st.get) if n else ""))
print(f"total samples {total}")
print("by stall:", ", ".join(f"{k[6:]} {100 * v / total:.1f}%" for k, v in by_stall.most_common(10)))
print("by opcode (samples %, executed):")
for op, n in by_op.most_common(18):
    print(f"  {op:10s} {100 * n / total:5.1f}%  {by_op_n[op]}")
print("hottest instructions:")
for n, idx, s, why in sorted(inst, reverse=True)[:top_n]:
    print(f"  {100 * n / total:5.2f}%  #{idx:5d}  {why[6:]:12s} {s}")
