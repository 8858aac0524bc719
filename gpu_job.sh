mkdir -p gpurun_out
OCTA_GROW_TIMING=1 timeout 300 python tools/grow_probe.py --batch 64 --check 24 2>&1 | grep -E "timing|rep 1|parity" | tail -4
