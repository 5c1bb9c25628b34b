timeout 1500 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider > gpurun_out/r2s_tests.log 2>&1; echo "exit $?"; tail -5 gpurun_out/r2s_tests.log
timeout 600 python bench.py --no-cpu --no-lib > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; echo "exit $?"; cut -c1-300 gpurun_out/r2s_bench.json
SFVOS_BF16_DGRAD=0 timeout 600 python bench.py --no-cpu --no-lib > gpurun_out/r2s_bench_f32dgrad.json 2> gpurun_out/r2s_bench.err; echo "exit $?"; cut -c1-300 gpurun_out/r2s_bench_f32dgrad.json
