mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_r01_v10.json 2> gpurun_out/bench_err.log; python -c "
import json; d=json.load(open('gpurun_out/bench_r01_v10.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'], d['cpu_baseline']['value'], d['config5_gan']['images_per_sec'], d['clocks'])"; tail -2 gpurun_out/bench_err.log
