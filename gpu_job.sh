mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pipeline_gpu.py tests/test_cli_gpu.py -m gpu -x -q > gpurun_out/pytest_cli.log 2>&1; tail -15 gpurun_out/pytest_cli.log
timeout 300 python tools/pipe_probe.py 10 8 32 0 2>&1 | grep "PROBE\|TRACE\|Error"
timeout 300 python tools/pipe_probe.py 12 8 32 1 2>&1 | grep "PROBE\|TRACE\|Error"
OCTA_EXTRA_SLOTS=24 timeout 300 python tools/pipe_probe.py 12 8 32 1 2>&1 | grep "PROBE\|TRACE\|Error"
timeout 300 python tools/pipe_probe.py 12 10 32 1 2>&1 | grep "PROBE\|TRACE\|Error"
