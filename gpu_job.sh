mkdir -p gpurun_out
for ty in 16 32; do
OCTA_VOX_TILE_Y=$ty timeout 600 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/bench_ty$ty.log 2>&1; tail -1 gpurun_out/bench_ty$ty.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print($ty, d['value'], d['config']['phase_ms'], d['roofline']['frac'])"
done
