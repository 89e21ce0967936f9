python bench.py --tc 2 --batch 235 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:spec_|y_producer|edge_epilogue|l0_combine|rowsum_planes" -s 13 -c 13 -o gpurun_out/prof_sp1 python bench.py --tc 2 --batch 235 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu4.log 2>&1
tail -3 gpurun_out/ncu4.log
