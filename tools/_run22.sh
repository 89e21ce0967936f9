timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
SNDVAE_STAGE_TIMING=1 timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/sp_bench15.json 2> gpurun_out/sp_bench15.err
tail -1 gpurun_out/sp_bench15.err | cut -c1-420; cut -c1-160 gpurun_out/sp_bench15.json
