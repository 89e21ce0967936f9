#!/bin/bash
# Round-end evidence on one B200: default bench line (+ stage timings), ncu launch list of one step of a 256-graph
# micro-batch, and an `ncu --set full` capture of the N^2-stage kernels of the same step.  Outputs in gpurun_out/.
mkdir -p gpurun_out
SNDVAE_STAGE_TIMING=1 timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
grep "sndvae stages" gpurun_out/bench_final.err | tail -3 > gpurun_out/bench_final.stages.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv \
  python bench.py --batch 256 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/launches.log 2>&1
timeout 900 ncu --set full --clock-control none -k "regex:spec_|y_producer|edge_epilogue|l0_combine|rowsum_planes" -s 14 -c 14 \
  -o gpurun_out/stage_full -f python bench.py --batch 256 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/stage_full.log 2>&1
ncu -i gpurun_out/stage_full.ncu-rep --page raw --csv > gpurun_out/stage_full_raw.csv 2>/dev/null
rm -f gpurun_out/stage_full.ncu-rep
cat gpurun_out/bench_final.json
# optional (about 9 GPU-minutes): speed-of-light metrics of EVERY launch of a step -> profiles/sol_<tag>_step_b256.txt via tools/sol_summary.py
if [ -n "$SOL_STEP" ]; then
  timeout 1000 ncu --metrics gpu__time_duration.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active \
    --clock-control none -c 540 --csv --log-file gpurun_out/sol_step.csv python bench.py --batch 256 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/sol_step.log 2>&1
fi

