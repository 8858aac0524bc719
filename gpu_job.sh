mkdir -p gpurun_out
nproc
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_b.json 2> gpurun_out/bench_r02_b.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_r02_b.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['ms_per_launch'], d['cpu_baseline']['value'], d['config5_gan']['images_per_sec'])
"; tail -3 gpurun_out/bench_r02_b.err
timeout 300 python tools/post_only.py 5 all 2>&1 | tail -3
