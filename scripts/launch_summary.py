"""Summarise an ncu --csv launch list (gpu__time_duration.sum): total time and count per kernel name."""
import csv, sys, collections
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
ui = hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    if r[ui] == "ns": v /= 1e3
    elif r[ui] == "ms": v *= 1e3
    tot[r[ki][:90]] += v; cnt[r[ki][:90]] += 1
allt = sum(tot.values())
for k, v in tot.most_common(30):
    print(f"{v/1e3:10.3f} ms {100*v/allt:5.1f}%  x{cnt[k]:4d}  {v/cnt[k]:9.1f} us/launch  {k}")
