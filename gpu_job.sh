timeout 900 python -m pytest tests/test_voxelize_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python tools/post_only.py 5 vox 2>&1 | tail -1
OCTA_VOX_TILE_Y=16 timeout 300 python tools/post_only.py 5 vox 2>&1 | tail -1
timeout 300 python tools/post_only.py 5 vox 2>&1 | tail -1
