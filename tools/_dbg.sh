timeout 300 python tools/bench_roi.py 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_roi_mask.py tests/test_gpu_kernels.py tests/test_gpu_fullsize.py -m gpu -x -q --no-header -p no:cacheprovider 2>&1 | tail -2
