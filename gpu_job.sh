mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_voxelize_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/post_only.py 5 vox 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vox_col_kernel -s 1 -c 1 -f -o gpurun_out/r02_vox_col_v3 python tools/post_only.py 1 vox > gpurun_out/ncu_vox.log 2>&1; tail -2 gpurun_out/ncu_vox.log
