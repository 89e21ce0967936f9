python tools/gpu_check.py --n 25 --b 7 --s 10 --tc 2 --chunk 3 > gpurun_out/sp_n25.log 2>&1; grep -E "RESULT|!!" gpurun_out/sp_n25.log
python tools/gpu_check.py --n 7 --b 5 --s 2 --tc 2 --chunk 2 > gpurun_out/sp_n7.log 2>&1; grep -E "RESULT|!!" gpurun_out/sp_n7.log
timeout 600 python tools/tc_vs_simt.py --n 256 --b 4 --s 2 --chunk 3 --tc 2 > gpurun_out/sp_n256.log 2>&1; grep -E "logits|O12|grad" gpurun_out/sp_n256.log | head -6
SNDVAE_STAGE_TIMING=1 python bench.py --tc 2 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/sp_bench3.json 2> gpurun_out/sp_bench3.err
tail -1 gpurun_out/sp_bench3.err; cut -c1-200 gpurun_out/sp_bench3.json
