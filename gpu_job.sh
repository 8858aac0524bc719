mkdir -p gpurun_out
timeout 900 python tools/grow_probe.py --batch 64 --check 24 --reps 2 > gpurun_out/grow_probe.log 2>&1
head -3 gpurun_out/grow_probe.log; tail -1 gpurun_out/grow_probe.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-1000
