timeout 900 python -m pytest tests/test_growth_gpu.py -m gpu -x -q 2>&1 | tail -2
OCTA_GROW_TIMING=1 timeout 300 python tools/grow_probe.py --batch 64 --reps 2 2>&1 | grep "timing" | tail -2 | cut -c1-330
P="timeout 300 python tools/pipe_probe.py 20 7 64 1"
$P 2>&1 | grep "PROBE\|Error"
$P 2>&1 | grep "PROBE\|Error"
