mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gan_gpu.py -q -x -s > gpurun_out/pytest_gan.log 2>&1; tail -6 gpurun_out/pytest_gan.log
timeout 200 python tools/gan_probe.py --batch 32 --reps 5 2>&1 | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/gan_launches_v4.csv python tools/gan_probe.py --batch 32 --reps 1 > gpurun_out/gan_ncu.log 2>&1; tail -1 gpurun_out/gan_ncu.log
