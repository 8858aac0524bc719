timeout 900 python -m pytest tests/test_csv_device_gpu.py tests/test_pipeline_gpu.py tests/test_cli_gpu.py -m gpu -x -q 2>&1 | tail -4
P="timeout 300 python tools/pipe_probe.py 20 7 64 1"
$P 2>&1 | grep "PROBE\|Error"
OCTA_CSV=host $P 2>&1 | grep "PROBE\|Error"
taskset -c 0-3 $P 2>&1 | grep "PROBE\|Error"
OCTA_CSV=host taskset -c 0-3 $P 2>&1 | grep "PROBE\|Error"
