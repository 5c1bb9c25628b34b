mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/t_kernels.log 2>&1; echo "kernels exit $?"; tail -15 gpurun_out/t_kernels.log
for G in 1 2 3; do for LP in 16 10; do SFVOS_TSTACK_G=$G SFVOS_TSTACK_LP=$LP timeout 120 python tools/bench_conv.py fast1 2>&1 | tail -1; done; done
SFVOS_TSTACK=0 timeout 120 python tools/bench_conv.py fast1 fast2 fast3 fast2+d fast3+d 2>&1 | tail -5
timeout 120 python tools/bench_conv.py fast2 fast3 fast2+d fast3+d 2>&1 | tail -4
SFVOS_TSTACK_LP=10 timeout 120 python tools/bench_conv.py fast2 fast3 fast2+d fast3+d 2>&1 | tail -4
