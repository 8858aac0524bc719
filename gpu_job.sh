OCTA_GROW_HOST_TIMING=1 timeout 600 python tools/step_probe.py 2>&1 | grep "host\]" | tail -3
