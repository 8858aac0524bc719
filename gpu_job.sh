mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log > gpurun_out/bench_line.json; python -c "
import json; d=json.load(open('gpurun_out/bench_line.json')); print(d['value'], d['e2e']['value'], d['config']['phase_ms'], d['roofline']['frac'], d['gpu_launches'], d['cpu_baseline']['value'])" || tail -20 gpurun_out/bench.log
