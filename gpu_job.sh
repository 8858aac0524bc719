free -g | head -2
timeout 600 python bench.py --steps 12 --warmup 4 --in-flight 3 --d2h-volume --no-gan --no-cpu-baseline > gpurun_out/bench_r02_d2hvol.json 2> gpurun_out/bench_r02_d2hvol.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_r02_d2hvol.json').read().strip().splitlines()[-1])
print('d2h-volume', d['value'], d['e2e'])
"; tail -3 gpurun_out/bench_r02_d2hvol.err
