mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -p no:cacheprovider -x -k "conv and umma" > gpurun_out/t_kernels.log 2>&1; echo "kernels exit $?"; tail -15 gpurun_out/t_kernels.log
timeout 120 python tools/bench_conv.py slow1 slow3 slow2+d 2>&1 | tail -3
SFVOS_PAIR=0 timeout 120 python tools/bench_conv.py slow1 slow3 slow2+d 2>&1 | tail -3
