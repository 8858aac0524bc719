mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pipeline_gpu.py -m gpu -x -q > gpurun_out/pytest_pipe.log 2>&1; tail -4 gpurun_out/pytest_pipe.log
for f in 1 2 3; do
timeout 900 python bench.py --no-cpu-baseline --steps 6 --in-flight $f > gpurun_out/bench_f$f.log 2>&1; tail -1 gpurun_out/bench_f$f.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print($f, d['value'], d['e2e']['value'], d['config']['phase_ms'])" || tail -5 gpurun_out/bench_f$f.log
done
