mkdir -p gpurun_out
nproc; free -g | sed -n 2p
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_r02_n8.json 2> gpurun_out/n8.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_r02_n8.json').read().strip().splitlines()[-1]); print('N8', d['value'], d['e2e']['value'], d['ms_per_step'])"; tail -3 gpurun_out/n8.err
OCTA_NO_PIN=1 timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r02_n8_nopin.json 2> gpurun_out/n8b.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_r02_n8_nopin.json').read().strip().splitlines()[-1]); print('N8 nopin', d['value'], d['e2e']['value'], d['ms_per_step'])"; tail -3 gpurun_out/n8b.err
timeout 600 $TR bench.py --gpus 8 --config 3 > gpurun_out/bench_r02_cfg3_n8.json 2> gpurun_out/cfg3n8.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_r02_cfg3_n8.json').read().strip().splitlines()[-1]); print('cfg3 N8', d['value'], d['files_only_s_max'], d['csv_set_sha256'])"; tail -3 gpurun_out/cfg3n8.err
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-gan > gpurun_out/bench_r02_n1_on8box.json 2> gpurun_out/n1b.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_r02_n1_on8box.json').read().strip().splitlines()[-1]); print('N1 same box', d['value'], d['e2e']['value'])"
