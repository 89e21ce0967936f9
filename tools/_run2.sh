SNDVAE_STAGE_TIMING=1 python bench.py --tc 2 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/sp_bench1.json 2> gpurun_out/sp_bench1.err
tail -3 gpurun_out/sp_bench1.err; cat gpurun_out/sp_bench1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/sp_launches.csv python bench.py --tc 2 --batch 235 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/sp_ncu.log 2>&1
