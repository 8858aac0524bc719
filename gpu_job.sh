mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gan_gpu.py -m gpu -x -q -s -k shipped 2>&1 | grep "shipped checkpoint:\|passed\|failed" | head -5
timeout 600 python bench.py --config 3 > gpurun_out/bench_r02_cfg3_n1.json 2> gpurun_out/cfg3.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_r02_cfg3_n1.json').read().strip().splitlines()[-1]); print(d['value'], d['files_only_s_max'], d['population'], d['csv_set_sha256'])"; tail -3 gpurun_out/cfg3.err
timeout 900 python -m pytest tests/test_cli_gpu.py tests/test_pipeline_gpu.py tests/test_gan_gpu.py -m gpu -x -q 2>&1 | tail -3
