mkdir -p gpurun_out
bash tools/gpu_tests.sh test_gpu_kernels test_gpu_slowfast test_gpu_roi_mask > gpurun_out/tests.log 2>&1; grep -E "^===|^exit|passed|failed|Error|error" gpurun_out/tests.log | head -30
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_v10.json 2> gpurun_out/bench_v10.err; echo "bench exit $?"; cat gpurun_out/bench_v10.json; tail -3 gpurun_out/bench_v10.err
python tools/profile_step.py --rows 30 > gpurun_out/prof_step_v10.txt 2>&1; echo "prof exit $?"; head -42 gpurun_out/prof_step_v10.txt | cut -c1-75,120-230
