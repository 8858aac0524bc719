timeout 900 python -m pytest tests/test_gan_gpu.py -m gpu -x -q -s 2>&1 | grep "shipped checkpoint:\|generator seed\|full size\|passed\|failed\|Error" | head
timeout 300 python tools/gan_probe.py --batch 32 --reps 3 2>&1 | tail -3
