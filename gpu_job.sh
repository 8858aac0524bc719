mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vox_col_kernel -s 1 -c 1 -f -o gpurun_out/r02_vox_col_v5 python tools/post_only.py 1 vox > gpurun_out/ncu_vox.log 2>&1; tail -1 gpurun_out/ncu_vox.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_h.json 2> gpurun_out/bench_r02_h.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_r02_h.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['ms_per_launch'], d['cpu_baseline']['value'], d['config5_gan']['images_per_sec'])
"; tail -3 gpurun_out/bench_r02_h.err
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -1
