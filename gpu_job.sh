mkdir -p gpurun_out
run() {
env $2 timeout 300 python bench.py --no-cpu-baseline --no-gan $3 > gpurun_out/exp_c.json 2> gpurun_out/exp_c.err
python -c "
import json
d=json.load(open('gpurun_out/exp_c.json')); print('$1', 'value %.1f e2e %.1f step %.1f ms loop %.1f ms'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['config']['phase_ms']['growth_loop_device']))" || tail -3 gpurun_out/exp_c.err
}
run k16_post2 OCTA_POST_WORKERS=2 ""
run k16_post1 OCTA_POST_WORKERS=1 ""
run k4_post2 OCTA_POST_WORKERS=2 "--steps 4"
run k16_post3 OCTA_POST_WORKERS=3 ""
