timeout 300 python -m pytest tests/test_pipeline_gpu.py -x -q 2>&1 | tail -1
K=10 OCTA_EXTRA_SLOTS=0 timeout 300 python tools/e2e_probe.py 2>&1 | tail -4 | sed 's/^/slots+0 /'
K=10 OCTA_EXTRA_SLOTS=3 timeout 300 python tools/e2e_probe.py 2>&1 | tail -4 | sed 's/^/slots+3 /'
