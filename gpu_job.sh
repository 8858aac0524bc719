mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_r01_v7.json 2> gpurun_out/bench_err.log; tail -c 1500 gpurun_out/bench_r01_v7.json; tail -3 gpurun_out/bench_err.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gan_conv3 -s 3 -c 1 -o gpurun_out/r01_prof_gan_conv3 -f python tools/gan_probe.py --batch 32 --reps 1 > gpurun_out/ncu_gan_full.log 2>&1; tail -2 gpurun_out/ncu_gan_full.log
