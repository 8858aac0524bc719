mkdir -p gpurun_out
# (1) every launch of one bench run (2 steps) with device times
ncu --metrics gpu__time_duration.sum --clock-control none -s 8700 -c 9100 --csv --log-file gpurun_out/r01_bench_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
# (2) full capture of the voxel tile kernel (batch 64) and of two mid-run k_commit / k_kill / k_eval launches
ncu --set full --clock-control none --import-source on -k regex:vox_tile -s 3 -c 1 -o gpurun_out/r01_prof_vox_tile python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_commit -s 300 -c 2 -o gpurun_out/r01_prof_k_commit python tools/grow_probe.py --batch 64 --reps 1 > gpurun_out/ncu_full2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_kill -s 300 -c 2 -o gpurun_out/r01_prof_k_kill python tools/grow_probe.py --batch 64 --reps 1 > gpurun_out/ncu_full3.log 2>&1
ls -la gpurun_out | tail -8
