"""Summarise an `ncu --page raw --csv` export: one line per profiled launch."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[0]
want = [("Kernel Name", "kernel", 30), ("gpu__time_duration.sum", "time", 10), ("dram__bytes_read.sum", "dram_rd", 10), ("dram__bytes_write.sum", "dram_wr", 10),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 6), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%", 6),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%", 7), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", 6),
        ("launch__registers_per_thread", "regs", 5), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%", 6),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1%", 6), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%", 6),
        ("smsp__issue_active.avg.pct", "issue%", 6), ("launch__grid_size", "grid", 8)]
idx = [(h.index(k) if k in h else -1, n, w) for k, n, w in want]
units = rows[1]
print(" ".join(n.ljust(w) for _, n, w in idx))
for r in rows[2:]:
    out = []
    for i, n, w in idx:
        v = r[i] if i >= 0 else "-"
        if i >= 0 and units[i]: v = v + units[i][:2] if n in ("time", "dram_rd", "dram_wr") else v
        out.append(v[:w].ljust(w))
    print(" ".join(out))
