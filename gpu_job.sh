mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gan_gpu.py -q -x -s > gpurun_out/pytest_gan.log 2>&1; tail -7 gpurun_out/pytest_gan.log
timeout 200 python tools/gan_probe.py --batch 32 --reps 5 2>&1 | tail -2
