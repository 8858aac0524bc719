mkdir -p gpurun_out
python tools/step_probe.py > gpurun_out/step_probe.log 2>&1; tail -12 gpurun_out/step_probe.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/grow_launches.csv python tools/grow_probe.py --batch 64 --reps 1 > gpurun_out/grow_ncu.log 2>&1
