"""Hottest SASS instructions + stall-reason totals from `ncu --page source --csv --print-source sass`."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[hi + 1:]:
    if r and r[0] == "Address":
        break                      # next launch
    if len(r) == len(hdr) and r[0].startswith("0x"):
        data.append(r)
S = idx['# Samples']
tot = sum(int(r[S]) for r in data)
print("total samples", tot, "instructions", len(data))
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(int(r[idx[h]]) for r in data) for h in stall_cols}
for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]:
    print(f"  {h:24s} {v:8d} {100 * v / tot:5.1f}%")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
for r in sorted(data, key=lambda r: -int(r[S]))[:n]:
    st = sorted([(int(r[idx[h]]), h[6:]) for h in stall_cols], reverse=True)[:2]
    print(f"{int(r[S]):7d} {100 * int(r[S]) / tot:5.1f}%  {r[idx['Source']].strip()[:64]:64s} {st}")
