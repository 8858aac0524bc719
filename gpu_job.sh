# final check of the round (the GPU budget left 2 minutes: smoke, a short bench, then the GPU suite one file per worker)
mkdir -p gpurun_out
timeout 25 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 40 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-gan > gpurun_out/bench_r02_final_short.json 2> gpurun_out/bench_r02_final_short.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_r02_final_short.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'])
"; tail -2 gpurun_out/bench_r02_final_short.err
timeout 70 python -m pytest tests -m gpu -q -n 4 --dist loadfile > gpurun_out/pytest_gpu_final.log 2>&1; tail -4 gpurun_out/pytest_gpu_final.log
