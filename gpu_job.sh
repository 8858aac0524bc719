mkdir -p gpurun_out
nproc; free -g | head -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 4 --warmup 3 > gpurun_out/bench4.log 2>&1; tail -1 gpurun_out/bench4.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('N=4', d['value'], d['e2e']['value'], d['config']['phase_ms'], d['ms_per_step'])" || tail -20 gpurun_out/bench4.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 4 --steps 1 --warmup 0 2>&1 | tail -2 | cut -c1-300
