#!/bin/bash
# Round 2, GPU session A: full GPU suite, ReLU-flip diagnosis, 20x repeat of the suite, compute-sanitizer logs.
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "=== full suite" ; timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/r2a_suite.log 2>&1; echo "exit $?"; tail -25 gpurun_out/r2a_suite.log
cp gpurun_out/parity_report.jsonl gpurun_out/r2a_parity_report.jsonl 2>/dev/null
echo "=== relu flip report"; timeout 600 python tools/relu_flip_report.py > gpurun_out/r2a_relu_flip.log 2>&1; echo "exit $?"; cat gpurun_out/r2a_relu_flip.log | tail -40
echo "=== 20x suite"; : > gpurun_out/r2a_loop20.log
for i in $(seq 1 20); do timeout 600 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider 2>&1 | tail -1 | sed "s/^/run $i: /" >> gpurun_out/r2a_loop20.log; done
cat gpurun_out/r2a_loop20.log
echo "=== sanitizer"
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 30 python -m pytest tests/test_gpu_slowfast.py -q --no-header -p no:cacheprovider -k "golden and 3-7 or reproducible and 3-7 or eval_mode_backward or fuse" > gpurun_out/r2a_sanitizer_$tool.log 2>&1
  echo "$tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r2a_sanitizer_$tool.log | tail -3
done
