python tools/gpu_check.py --n 25 --b 7 --s 10 --tc 2 --chunk 3 > gpurun_out/sp_n25.log 2>&1; grep -E "RESULT|!!" gpurun_out/sp_n25.log
timeout 600 python tools/tc_vs_simt.py --n 256 --b 4 --s 2 --chunk 3 --tc 2 > gpurun_out/sp_n256.log 2>&1; grep -E "logits|O12|e1_deconv|grad" gpurun_out/sp_n256.log | head
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/sp_launches2.csv python bench.py --tc 2 --batch 235 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/sp_ncu2.log 2>&1
SNDVAE_FFT_THREADS=800 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/sp_launches3.csv python bench.py --tc 2 --batch 235 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/sp_ncu3.log 2>&1
SNDVAE_STAGE_TIMING=1 python bench.py --tc 2 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/sp_bench2.json 2> gpurun_out/sp_bench2.err
tail -1 gpurun_out/sp_bench2.err; cut -c1-400 gpurun_out/sp_bench2.json
