timeout 900 python -m pytest tests/test_csv_device_gpu.py tests/test_pipeline_gpu.py tests/test_cli_gpu.py -m gpu -x -q 2>&1 | tail -3
P="timeout 300 python tools/pipe_probe.py 20 7 64 1"
OCTA_CSV=device $P 2>&1 | grep "PROBE\|Error"
OCTA_CSV=host $P 2>&1 | grep "PROBE\|Error"
OCTA_CSV=device $P 2>&1 | grep "PROBE\|Error"
OCTA_CSV=host $P 2>&1 | grep "PROBE\|Error"
