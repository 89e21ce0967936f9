timeout 900 python -m pytest tests -m gpu -x -q -k "n100 or spectral or parity" > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py --nodes 100 --batch 1024 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n100b.json 2> gpurun_out/bench_n100b.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n100b.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
