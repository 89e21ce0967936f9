timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
SNDVAE_STAGE_TIMING=1 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/sp_bench5.json 2> gpurun_out/sp_bench5.err
tail -1 gpurun_out/sp_bench5.err; cut -c1-200 gpurun_out/sp_bench5.json
SNDVAE_CUBLAS_PEDANTIC=1 SNDVAE_STAGE_TIMING=1 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/sp_bench5p.json 2> gpurun_out/sp_bench5p.err
tail -1 gpurun_out/sp_bench5p.err; cut -c1-200 gpurun_out/sp_bench5p.json
