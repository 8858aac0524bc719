# scratch script for `gpurun -- 'bash gpu_job.sh'`
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
nproc
timeout 900 python -m pytest tests/test_growth_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q > gpurun_out/pytest_growth.log 2>&1; tail -5 gpurun_out/pytest_growth.log
OCTA_GROW_GRAPH=0 OCTA_BALL_ORDER=always timeout 300 python tools/pipe_probe.py 10 8 2>&1 | grep PROBE
OCTA_GROW_GRAPH=1 OCTA_BALL_ORDER=always timeout 300 python tools/pipe_probe.py 10 8 2>&1 | grep PROBE
OCTA_GROW_GRAPH=0 timeout 300 python tools/pipe_probe.py 10 8 2>&1 | grep PROBE
timeout 300 python tools/pipe_probe.py 10 8 2>&1 | grep PROBE
timeout 300 python tools/pipe_probe.py 10 12 2>&1 | grep PROBE
timeout 300 python tools/pipe_probe.py 10 16 2>&1 | grep PROBE
timeout 300 python tools/pipe_probe.py 10 12 16 2>&1 | grep PROBE
timeout 300 python tools/pipe_probe.py 10 8 32 1 2>&1 | grep PROBE
timeout 300 python tools/pipe_probe.py 10 12 32 1 2>&1 | grep PROBE
OCTA_GROW_TIMING=1 timeout 300 python tools/grow_probe.py --batch 32 --reps 2 2>&1 | tail -8
