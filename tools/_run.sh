mkdir -p gpurun_out
bash tools/gpu_tests.sh test_gpu_model test_gpu_slowfast test_gpu_kernels > gpurun_out/tests.log 2>&1; grep -E "^===|^exit|passed|failed|Error|error|assert" gpurun_out/tests.log | head -30
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_v12.json 2> gpurun_out/bench_v12.err; echo "bench exit $?"; tail -5 gpurun_out/bench_v12.err; python -c "
import json; d=json.load(open('gpurun_out/bench_v12.json')); print(d['value'], d['ms_per_step'], d['config']['launch'], d['config']['eager_ms_per_step'], d['e2e']['value']); print({k:(v['tflops'],v['ms_per_step']) for k,v in d['roofline']['per_kernel'].items()}); print(d['roofline']['all_tensor_kernels'])"
python tools/profile_step.py --rows 22 > gpurun_out/prof_step_v12.txt 2>&1; head -34 gpurun_out/prof_step_v12.txt | cut -c1-75,120-230
