mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -p no:cacheprovider -k "wgrad" > gpurun_out/t_kernels.log 2>&1; echo "kernels exit $?"; tail -15 gpurun_out/t_kernels.log
timeout 120 python tools/bench_conv.py fast1+w fast2+w fast3+w slow1+w slow3+w f2s1+w 2>&1 | tail -6
SFVOS_WGRAD_C32=0 timeout 120 python tools/bench_conv.py fast2+w fast3+w 2>&1 | tail -2
