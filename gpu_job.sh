mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_growth_gpu.py -x -q > gpurun_out/pytest_growth.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_growth.log
tail -40 gpurun_out/pytest_growth.log
