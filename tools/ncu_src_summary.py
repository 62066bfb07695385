"""Summarise an `ncu --page source --csv` export: stall reasons by executed-count bucket and the hottest SASS lines."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = rows[1]
data = rows[2:]
iS, iSrc, iEx = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
stalls = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[iS]) for r in data)
print("kernel:", rows[0][1][:100])
print("total samples", tot, "SASS instructions", len(data))
byex = collections.defaultdict(lambda: [0, 0, collections.Counter()])
for r in data:
    b = byex[int(r[iEx])]
    b[0] += 1
    b[1] += int(r[iS])
    for i in stalls:
        b[2][hdr[i]] += int(r[i])
for k, (n, s, c) in sorted(byex.items(), key=lambda x: -x[1][1])[:8]:
    print(f"exec={k:8d} n_instr={n:5d} samples={s:6d}  ", [(a[6:], b) for a, b in c.most_common(6)])
print("hottest lines:")
for k in sorted(sorted(range(len(data)), key=lambda k: -int(data[k][iS]))[:top_n]):
    r = data[k]
    st = sorted(((int(r[i]), hdr[i][6:]) for i in stalls), reverse=True)[:2]
    print(f"{k:5d} {r[iS]:>6s} {r[iEx]:>8s}  {r[iSrc].strip()[:78]:78s} {st}")
