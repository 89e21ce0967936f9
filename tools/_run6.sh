timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
SNDVAE_STAGE_TIMING=1 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/sp_bench4.json 2> gpurun_out/sp_bench4.err
tail -1 gpurun_out/sp_bench4.err; cat gpurun_out/sp_bench4.json
