set -x
python bench.py --batch 235 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/plain_final.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_final.csv python bench.py --batch 235 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_final_l.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:spec_|y_producer|edge_epilogue|l0_combine|rowsum_planes" -s 14 -c 14 -o gpurun_out/prof_final python bench.py --batch 235 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_final_f.log 2>&1
tail -2 gpurun_out/ncu_final_f.log | cut -c1-300
