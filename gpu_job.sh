mkdir -p gpurun_out
python tools/grow_probe.py --batch 64 --check 2 --reps 2 > gpurun_out/grow_probe.log 2>&1
cat gpurun_out/grow_probe.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-1500
ncu --set full --clock-control none --import-source on -k regex:vox_tile -s 1 -c 1 -o gpurun_out/prof_vox2 python bench.py --steps 1 --warmup 1 --batch 8 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
