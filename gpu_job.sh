mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_raster2d_gpu.py tests/test_pipeline_gpu.py -x -q > gpurun_out/pytest_r2d.log 2>&1; tail -2 gpurun_out/pytest_r2d.log
run() {
env $2 timeout 300 python bench.py --steps 8 --warmup 4 --no-cpu-baseline --no-gan $3 > gpurun_out/exp_c.json 2> gpurun_out/exp_c.err
python -c "
import json
d=json.load(open('gpurun_out/exp_c.json')); print('$1', 'value %.1f e2e %.1f step %.1f ms loop %.1f ms'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['config']['phase_ms']['growth_loop_device']))" || tail -3 gpurun_out/exp_c.err
}
run r2d_6x32 A=1 "--in-flight 6"
run r2d_8x32 A=1 "--in-flight 8"
run r2d_6x32_again A=1 "--in-flight 6"
