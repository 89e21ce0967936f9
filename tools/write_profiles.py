"""Turn the raw outputs of tools/final_profile.sh (gpurun_out/) into the tracked summaries under profiles/:
   ncu_<tag>_final_b256.txt (ncu --set full, N^2-stage kernels), launches_<tag>_final_b256.txt (launch list of one step),
   bench_<tag>_n256_b4096_1gpu_final.json (+ .stages.txt); <tag> = $PROFILE_TAG (default r2)."""
import csv, os, re, shutil, subprocess, sys
TAG = os.environ.get("PROFILE_TAG", "r2")          # round tag in the output file names
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
rows = list(csv.reader(open(os.path.join(G, "stage_full_raw.csv"))))
h, units = rows[0], rows[1]
col = h.index


def val(r, name):
    i = col(name); v = float(r[i].replace(",", ""))
    return v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}.get(units[i], 1.0)


out = ["# ncu --set full --clock-control none, N=256, 256 graphs (one micro-batch), second step of: python bench.py --batch 256 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e",
       "# -k regex:spec_|y_producer|edge_epilogue|l0_combine|rowsum_planes -s 14 -c 14   (per launch, cold cache, serialised; units: ms / GB / %); tools/final_profile.sh",
       "kernel | time_ms | dram_rd_GB | dram_wr_GB | dram_% | sm_% | tensor_% | occ_% | regs | l2_% | l1_% | issue_% | warp_inst"]
tt = tb = 0.0
for r in rows[2:]:
    name = r[col("Kernel Name")].split("(")[0][:46]
    t, rd, wr = val(r, "gpu__time_duration.sum"), val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
    g = lambda n: float(r[col(n)].replace(",", "")) if n in h else float("nan")
    out.append(f"{name} | {t:.3f} | {rd:.3f} | {wr:.3f} | {g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
               f"{g('sm__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | "
               f"{g('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | {int(g('launch__registers_per_thread'))} | "
               f"{g('lts__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {g('l1tex__throughput.avg.pct_of_peak_sustained_active'):.1f} | "
               f"{g('sm__issue_active.avg.pct_of_peak_sustained_elapsed'):.1f} | {int(g('smsp__inst_executed.sum'))}")
    if re.search(r"spec_fft|spec_gemm_k|spec_wgrad_k", name):
        tt += t; tb += rd + wr
out.append(f"# spectral e2e layer-1 stage (7 launches): {tt:.2f} ms, DRAM traffic {tb:.1f} GB = {tb / 256 * 1000:.0f} MB per graph "
           "(compulsory: 344.8 MB per graph, dO counted once)")
import json
tpath = os.path.join(P, "traffic.json")      # read by bench.py (roofline.traffic): measured DRAM bytes of the stage per graph, keyed by N
try:
    tj = json.load(open(tpath))
except Exception:
    tj = {}
head_sha = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
tj["n256"] = {"bytes_per_graph": round(tb / 256 * 1e9), "stage_ms_per_256_graphs": round(tt, 3),
              "source": f"profiles/ncu_{TAG}_final_b256.txt: ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of the stage's 7 launches "
                        f"on a 256-graph micro-batch (tools/final_profile.sh at {head_sha})"}
json.dump(tj, open(tpath, "w"), indent=1)
open(os.path.join(P, f"ncu_{TAG}_final_b256.txt"), "w").write("\n".join(out) + "\n")
print(out[-1])

skip = int(sys.argv[1]) if len(sys.argv) > 1 else -1
txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_summary.py"), os.path.join(G, "launches.csv"), str(skip), "45"],
                     capture_output=True, text=True).stdout
tot = float(re.search(r"total ([\d.]+) ms", txt).group(1))
st = sum(float(m.group(1)) for m in re.finditer(r"^(?:void )?spec_(?:fft|gemm_k|wgrad_k)[^\n]*ms=\s*([\d.]+)", txt, re.M))
head = (f"# ncu --metrics gpu__time_duration.sum --clock-control none; second step (the launches between the first two tf_adam_k) of: python bench.py --batch 256 --steps 1 "
        "--warmup 1 --no-cpu-baseline --no-e2e (N=256, one micro-batch of 256 graphs; per-launch times are cold-cache and serialised)\n")
tail = f"# spectral e2e layer-1 stage (spec_fft_*, spec_gemm_*, spec_wgrad_k): {st:.3f} ms = {st / tot * 100:.1f}% of the launch-list time\n"
open(os.path.join(P, f"launches_{TAG}_final_b256.txt"), "w").write(head + txt + tail)
print(tail.strip())
shutil.copy(os.path.join(G, "bench_final.json"), os.path.join(P, f"bench_{TAG}_n256_b4096_1gpu_final.json"))
shutil.copy(os.path.join(G, "bench_final.stages.txt"), os.path.join(P, f"bench_{TAG}_n256_b4096_1gpu_final.stages.txt"))
