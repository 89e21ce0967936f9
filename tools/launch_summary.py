"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: count, total ms, share."""
import csv, collections, sys
rows = list(csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"')))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[skip:]:
    k = r["Kernel Name"].split("(")[0]
    agg[k][0] += 1; agg[k][1] += float(r["Metric Value"]) / 1e6
tot = sum(v[1] for v in agg.values())
print(f"launches {len(rows) - skip}  total {tot:.3f} ms")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]:
    print(f"{k[:70]:70s} n={v[0]:4d} ms={v[1]:9.3f} share={v[1] / tot * 100:5.1f}%")
