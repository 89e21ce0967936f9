"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: count, total ms, share.
usage: launch_summary.py launches.csv [skip | -1 = the step between the first two tf_adam_k launches] [top]"""
import csv, collections, sys
rows = list(csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"')))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
end = len(rows)
if skip < 0:      # auto: the launches between the first and the second Adam update = one whole step after the warm-up step
    adam = [i for i, r in enumerate(rows) if "tf_adam_k" in r["Kernel Name"]]
    skip, end = adam[0] + 1, adam[1] + 1
    rows = rows[:end]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[skip:]:
    k = r["Kernel Name"].split("(")[0]
    agg[k][0] += 1; agg[k][1] += float(r["Metric Value"]) / 1e6
tot = sum(v[1] for v in agg.values())
print(f"launches {len(rows) - skip}  total {tot:.3f} ms")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]:
    print(f"{k[:70]:70s} n={v[0]:4d} ms={v[1]:9.3f} share={v[1] / tot * 100:5.1f}%")
