timeout 900 python -m pytest tests/test_growth_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -2
OCTA_GROW_HOST_TIMING=1 timeout 300 python tools/pipe_probe.py 20 7 64 1 2>&1 | grep "octa grow host\|PROBE" | tail -3
taskset -c 0-3 timeout 300 python tools/pipe_probe.py 20 7 64 1 2>&1 | grep "PROBE\|Error"
taskset -c 0-3 timeout 300 python tools/pipe_probe.py 20 7 64 0 2>&1 | grep "PROBE\|Error"
