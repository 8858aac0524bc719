mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log > gpurun_out/bench_line.json; python -c "
import json; d=json.load(open('gpurun_out/bench_line.json')); print(d['value'], d['e2e']['value'], d['config']['phase_ms'], d['roofline']['frac'], d['gpu_launches'], d['cpu_baseline']['value'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 4200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; tail -2 gpurun_out/ncu_launch.log | cut -c1-200; wc -l gpurun_out/launches.csv
