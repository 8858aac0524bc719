mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_i.json 2> gpurun_out/bench_r02_i.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_r02_i.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['ms_per_launch'], d['cpu_baseline']['value'], d['config5_gan']['images_per_sec'])
"; tail -3 gpurun_out/bench_r02_i.err
