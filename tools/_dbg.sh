timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_roi_mask.py -m gpu -x -q --no-header -p no:cacheprovider > gpurun_out/r2u_tests.log 2>&1; echo "exit $?"; tail -5 gpurun_out/r2u_tests.log
timeout 600 python tools/profile_timeline.py --out gpurun_out/r2u_timeline.csv 2>&1 | tail -36
timeout 600 python bench.py --no-cpu --no-lib > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; echo "exit $?"; cut -c1-200 gpurun_out/r2u_bench.json
timeout 300 python tools/bench_l2.py 2>&1 | tee gpurun_out/r2u_l2.json
