mkdir -p gpurun_out
bash tools/gpu_tests.sh test_gpu_kernels test_gpu_slowfast test_gpu_roi_mask > gpurun_out/tests.log 2>&1; grep -E "^===|^exit|passed|failed|Error|error" gpurun_out/tests.log | head -30
timeout 120 python tools/bench_conv.py f2s1+d f2s2+d 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_v8.json 2> gpurun_out/bench_v8.err; echo "bench exit $?"; cat gpurun_out/bench_v8.json; tail -3 gpurun_out/bench_v8.err
python tools/profile_step.py --rows 24 > gpurun_out/prof_step_v8.txt 2>&1; echo "prof exit $?"; head -36 gpurun_out/prof_step_v8.txt | cut -c1-75,120-230
