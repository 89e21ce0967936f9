#!/bin/bash
# A/B harness for the kernels of the N^2 stage on one 256-graph micro-batch at N=256: per-launch time and DRAM bytes of
# y_producer_tc_k, the transforms, the per-frequency GEMMs and spec_wgrad_k from an ncu launch list (cold cache, serialised).
# Each argument "<fwd> <inv>" is one run with SNDVAE_FFT_ORDER / SNDVAE_FFT_ORDER_INV set (-1 = default); any other
# diagnostic switch of DESIGN.md is taken from the caller's environment, e.g.
#   SNDVAE_FFT_BULK_INV=1 bash tools/ab_fft_order.sh "-1 -1"      vs      bash tools/ab_fft_order.sh "-1 -1"
mkdir -p gpurun_out
for cfg in "$@"; do
  set -- $cfg
  tag="f${1}_i${2}"
  SNDVAE_FFT_ORDER=$1 SNDVAE_FFT_ORDER_INV=$2 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -k "regex:spec_fft|spec_wgrad_k|spec_gemm_k|y_producer_tc_k" --csv --log-file gpurun_out/ab_${tag}.csv \
    python bench.py --batch 256 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ab_${tag}.log 2>&1
  echo "== order fwd=$1 inv=$2 (exit $?)"
  python - gpurun_out/ab_${tag}.csv <<'EOF'
import csv, sys, collections
rows = list(csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"')))
per = collections.OrderedDict()
for r in rows:
    key = (r["ID"], r["Kernel Name"].split("(")[0][-44:])
    per.setdefault(key, {})[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
def to(v, u, base):
    f = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}
    return v * f.get(u, 1.0)
keys = list(per)
for k in keys[-9:]:                      # the last step: y_producer x2, fwd Y, gemm, inv O, fwd dO, gemm, inv dY, wgrad
    m = per[k]
    t = to(*m["gpu__time_duration.sum"], "ms"); rd = to(*m["dram__bytes_read.sum"], "GB"); wr = to(*m["dram__bytes_write.sum"], "GB")
    print(f"  {k[1]:46s} {t:7.3f} ms  rd {rd:6.2f} GB  wr {wr:6.2f} GB  {(rd + wr) / t:6.2f} TB/s")
EOF
done
