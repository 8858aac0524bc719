mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
OCTA_GROW_TIMING=1 timeout 300 python tools/grow_probe.py --batch 64 --check 24 2>&1 | grep -E "timing|rep 1|parity" | tail -4
OCTA_GROW_TIMING=1 timeout 300 python tools/grow_probe.py --batch 1 2>&1 | grep -E "timing|rep 1" | tail -3
