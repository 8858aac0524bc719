# scratch script for `gpurun -- 'bash gpu_job.sh'`: full verification of the current build on one B200
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_line.json 2> gpurun_out/bench_err.log; tail -c 1200 gpurun_out/bench_line.json; tail -2 gpurun_out/bench_err.log
timeout 200 python tools/gan_probe.py --batch 32 --reps 5 2>&1 | tail -2
timeout 300 python tools/e2e_probe.py 2>&1 | tail -4
