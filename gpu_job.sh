mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_raster2d_gpu.py -m gpu -x -q -s > gpurun_out/pytest_r2d.log 2>&1; tail -25 gpurun_out/pytest_r2d.log
timeout 600 python -m pytest tests/test_pipeline_gpu.py tests/test_voxelize_gpu.py -m gpu -x -q > gpurun_out/pytest_pipe.log 2>&1; tail -5 gpurun_out/pytest_pipe.log
timeout 300 python tools/pipe_probe.py 8 8 2>&1 | grep PROBE
timeout 300 python tools/step_probe.py 2>&1 | tail -12
