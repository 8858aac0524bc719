mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_growth_gpu.py -x -q > gpurun_out/pytest_growth.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_growth.log
tail -5 gpurun_out/pytest_growth.log
python tools/grow_probe.py --batch 64 --check 8 --reps 2 > gpurun_out/grow_probe.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3400 --csv --log-file gpurun_out/grow_launches.csv python tools/grow_probe.py --batch 64 --reps 1 > gpurun_out/grow_ncu.log 2>&1
cat gpurun_out/grow_probe.log
