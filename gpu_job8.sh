mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r02_n8_v2.json 2> gpurun_out/n8.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_r02_n8_v2.json').read().strip().splitlines()[-1]); print('N8', d['value'], d['e2e']['value'], d['ms_per_step'])"; tail -3 gpurun_out/n8.err
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-gan > gpurun_out/bench_r02_n1_on8box_v2.json 2> gpurun_out/n1b.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_r02_n1_on8box_v2.json').read().strip().splitlines()[-1]); print('N1 same box', d['value'], d['e2e']['value'])"
