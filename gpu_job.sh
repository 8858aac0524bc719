mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_growth_gpu.py -x -q > gpurun_out/pytest_growth.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_growth.log
tail -12 gpurun_out/pytest_growth.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench2.log 2>&1; tail -1 gpurun_out/bench2.log | cut -c1-700
