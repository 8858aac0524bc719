mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_growth_gpu.py tests/test_pipeline_gpu.py -x -q > gpurun_out/pytest_growth.log 2>&1; tail -2 gpurun_out/pytest_growth.log
for i in 1 2; do
timeout 300 python bench.py --steps 8 --warmup 4 --no-cpu-baseline --no-gan > gpurun_out/exp_regcap.json 2> gpurun_out/exp_regcap.err
python -c "
import json
d=json.load(open('gpurun_out/exp_regcap.json')); print('regcap 6x32', 'value %.1f e2e %.1f step %.1f ms loop %.1f ms launches %d'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['config']['phase_ms']['growth_loop_device'], d['gpu_launches'])); print(d['e2e'])" || tail -3 gpurun_out/exp_regcap.err
done
