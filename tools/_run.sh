mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -p no:cacheprovider -x -k "wgrad and umma" > gpurun_out/t_kernels.log 2>&1; echo "kernels exit $?"; tail -3 gpurun_out/t_kernels.log
timeout 120 python tools/bench_conv.py slow1+w slow3+w 2>&1 | tail -2
SFVOS_WGRAD_PAIR=0 timeout 120 python tools/bench_conv.py slow1+w slow3+w 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_roi_mask.py tests/test_gpu_slowfast.py -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -3
