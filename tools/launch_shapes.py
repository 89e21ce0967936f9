"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per (kernel, grid): count, total / mean ms.
usage: launch_shapes.py launches.csv [name-filter] [top]"""
import csv, collections, sys
rows = list(csv.DictReader(l for l in open(sys.argv[1], errors="replace") if l.startswith('"')))
flt = sys.argv[2] if len(sys.argv) > 2 else ""
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for r in rows:
    ms = float(r["Metric Value"]) / 1e6
    tot += ms
    k = r["Kernel Name"].split("(")[0]
    if flt and flt not in k:
        continue
    agg[(k[:60], r.get("Grid Size", ""))][0] += 1; agg[(k[:60], r.get("Grid Size", ""))][1] += ms
sel = sum(v[1] for v in agg.values())
print(f"launches {len(rows)} total {tot:.3f} ms; selected {sel:.3f} ms")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 40]:
    print(f"{k[0]:60s} grid={k[1]:>22s} n={v[0]:4d} ms={v[1]:9.3f} mean={v[1] / v[0]:8.4f}")
