mkdir -p gpurun_out
nproc
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --config 3 > gpurun_out/bench_r02_cfg3_n2.json 2> gpurun_out/cfg3n2.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_r02_cfg3_n2.json').read().strip().splitlines()[-1]); print(d['value'], d['files_only_s_max'], d['population']['edges_mean'], d['csv_set_sha256'], d['csv_files_match_gathered_tables'])"; tail -3 gpurun_out/cfg3n2.err
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r02_n2.json 2> gpurun_out/n2.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_r02_n2.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['ms_per_step'])"; tail -3 gpurun_out/n2.err
timeout 600 $TR bench.py --gpus 2 --config 5 --steps 4 > gpurun_out/bench_r02_cfg5_n2.json 2> gpurun_out/cfg5n2.err; tail -c 300 gpurun_out/bench_r02_cfg5_n2.json; tail -3 gpurun_out/cfg5n2.err
