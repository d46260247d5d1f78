"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, mean, total."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]; ki = h.index("Kernel Name"); vi = h.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    if len(r) > vi:
        agg[r[ki][:70]].append(float(r[vi].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:70s} n={len(v):5d} mean={sum(v)/len(v)/1e3:9.1f} us total={sum(v)/1e6:8.2f} ms {100*sum(v)/tot:5.1f}%")
