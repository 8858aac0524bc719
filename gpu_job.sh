mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kdorder_gpu.py -x -q > gpurun_out/pytest_kd.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_kd.log
tail -12 gpurun_out/pytest_kd.log
timeout 900 python tools/grow_probe.py --batch 64 --check 24 --reps 2 > gpurun_out/grow_probe.log 2>&1
head -4 gpurun_out/grow_probe.log; tail -1 gpurun_out/grow_probe.log
