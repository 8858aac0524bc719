mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_voxelize_gpu.py tests/test_pipeline_gpu.py tests/test_raster2d_gpu.py -m gpu -x -q > gpurun_out/pytest_post.log 2>&1; tail -5 gpurun_out/pytest_post.log
timeout 300 python tools/post_only.py 5 all 2>&1 | tail -4
OCTA_VOX_TILE_Y=16 timeout 300 python tools/post_only.py 5 vox 2>&1 | tail -1
OCTA_VOX_TILE_Y=4 timeout 300 python tools/post_only.py 5 vox 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vox_tile_kernel -s 1 -c 1 -f -o gpurun_out/r02_vox_v4 python tools/post_only.py 1 vox > gpurun_out/ncu_vox.log 2>&1; tail -2 gpurun_out/ncu_vox.log
timeout 900 python -m pytest tests/test_gan_gpu.py tests/test_growth_gpu.py -m gpu -x -q > gpurun_out/pytest_rest.log 2>&1; tail -5 gpurun_out/pytest_rest.log
