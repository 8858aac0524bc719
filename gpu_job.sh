mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_growth_gpu.py tests/test_voxelize_gpu.py -x -q > gpurun_out/pytest_gv.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gv.log
tail -5 gpurun_out/pytest_gv.log
python tools/grow_probe.py --batch 64 --check 4 --reps 2 > gpurun_out/grow_probe.log 2>&1
cat gpurun_out/grow_probe.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 4500 --csv --log-file gpurun_out/grow_launches.csv python tools/grow_probe.py --batch 64 --reps 1 > gpurun_out/grow_ncu.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log
