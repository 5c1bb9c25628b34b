mkdir -p gpurun_out
bash tools/gpu_tests.sh test_gpu_model test_gpu_roi_mask > gpurun_out/tests.log 2>&1; grep -E "^===|^exit|passed|failed|Error|error|assert" gpurun_out/tests.log | head -30
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_v11.json 2> gpurun_out/bench_v11.err; echo "bench exit $?"; cut -c1-200 gpurun_out/bench_v11.json; python -c "
import json; d=json.load(open('gpurun_out/bench_v11.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(json.dumps(d['roofline'], indent=0)[:1800])"
