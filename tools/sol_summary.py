"""Per-kernel speed-of-light table of one whole train step from an ncu metrics CSV (tools: see the header line it writes).
usage: python tools/sol_summary.py gpurun_out/sol_step.csv > profiles/sol_r2_step_b256.txt"""
import csv, collections, sys
rows = list(csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"')))
per = collections.OrderedDict()
for r in rows:
    per.setdefault(r["ID"], {"name": r["Kernel Name"].split("(")[0]})[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
ids = list(per)
adam = [i for i, k in enumerate(ids) if "tf_adam_k" in per[k]["name"]]
sel = ids[adam[0] + 1: adam[1] + 1]          # the launches between the first two Adam updates = one step after the warm-up step
f = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}
def val(m, k): v, u = m[k]; return v * f.get(u, 1.0)
agg = collections.OrderedDict()
for k in sel:
    m = per[k]; t = val(m, "gpu__time_duration.sum")
    a = agg.setdefault(m["name"], {"n": 0, "t": 0.0, "dram": 0.0, "sm": 0.0, "tc": 0.0, "gb": 0.0, "regs": 0, "occ": 0.0})
    a["n"] += 1; a["t"] += t
    a["dram"] += t * m["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"][0]
    a["sm"] += t * m["sm__throughput.avg.pct_of_peak_sustained_elapsed"][0]
    a["tc"] += t * m["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"][0]
    a["occ"] += t * m["sm__warps_active.avg.pct_of_peak_sustained_active"][0]
    a["gb"] += val(m, "dram__bytes_read.sum") + val(m, "dram__bytes_write.sum")
    a["regs"] = int(m["launch__registers_per_thread"][0])
tot = sum(a["t"] for a in agg.values())
print("# ncu --metrics gpu__time_duration,gpu__dram_throughput,sm__throughput,dram__bytes_{read,write},sm__pipe_tensor_cycles_active,launch__registers_per_thread,"
      "sm__warps_active --clock-control none; the launches between the first two tf_adam_k of: python bench.py --batch 256 --steps 1 --warmup 1 "
      "--no-cpu-baseline --no-e2e  (N=256, one micro-batch of 256 graphs; cold-cache, serialised launches; percentages are time-weighted means)")
print(f"# {len(sel)} launches, {tot:.3f} ms")
print("kernel | launches | ms | share_% | dram_% | sm_% | tensor_% | occ_% | regs | dram_GB | GB/s")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
    t = a["t"]
    if t / tot < 0.002: continue
    print(f"{name[:60]} | {a['n']} | {t:.3f} | {100 * t / tot:.1f} | {a['dram'] / t:.1f} | {a['sm'] / t:.1f} | {a['tc'] / t:.1f} | {a['occ'] / t:.1f} | {a['regs']} | {a['gb']:.2f} | {a['gb'] / t * 1e3:.0f}")
rest = sum(a["t"] for a in agg.values() if a["t"] / tot < 0.002)
print(f"# kernels below 0.2 % each: {rest:.3f} ms ({100 * rest / tot:.1f} %)")
