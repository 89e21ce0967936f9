timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
SNDVAE_STAGE_TIMING=1 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/sp_bench6.json 2> gpurun_out/sp_bench6.err
tail -1 gpurun_out/sp_bench6.err; cut -c1-200 gpurun_out/sp_bench6.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/sp_launches6.csv python bench.py --batch 235 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/sp_ncu6.log 2>&1
