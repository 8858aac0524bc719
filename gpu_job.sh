mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -2
for r in 1 2; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$r bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/bench2_$r.log 2>&1; tail -1 gpurun_out/bench2_$r.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('N=2', d['value'], d['e2e']['value'], d['config']['phase_ms'], d['ms_per_step'])" || tail -20 gpurun_out/bench2_$r.log
done
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('N=1', d['value'], d['e2e']['value'], d['config']['phase_ms'])" || tail -5 gpurun_out/bench.log
