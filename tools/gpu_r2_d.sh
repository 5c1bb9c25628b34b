#!/bin/bash
# Round 2, GPU session D: backbone/model tests, C4 pipeline (native FPN/RPN head on/off), ncu launch list of the bench step.
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
echo "=== tests"; timeout 1500 python -m pytest tests/test_gpu_backbone.py tests/test_gpu_model.py tests/test_gpu_sequence.py tests/test_gpu_callers.py tests/test_gpu_slowfast.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/r2d_tests.log 2>&1; echo "exit $?"; tail -15 gpurun_out/r2d_tests.log
cp gpurun_out/parity_report.jsonl gpurun_out/r2d_parity_report.jsonl 2>/dev/null
echo "=== pipeline C4, native FPN + RPN head"; timeout 900 python tools/bench_pipeline.py --sequences 4 --frames 24 > gpurun_out/r2d_pipeline_native.jsonl 2> gpurun_out/r2d_pipeline_native.err; echo "exit $?"; tail -3 gpurun_out/r2d_pipeline_native.err; cat gpurun_out/r2d_pipeline_native.jsonl | cut -c1-900
echo "=== pipeline C4, torchvision FPN + RPN head"; SFVOS_NATIVE_BACKBONE=0 timeout 900 python tools/bench_pipeline.py --sequences 4 --frames 24 > gpurun_out/r2d_pipeline_tv.jsonl 2> gpurun_out/r2d_pipeline_tv.err; echo "exit $?"; tail -3 gpurun_out/r2d_pipeline_tv.err; cat gpurun_out/r2d_pipeline_tv.jsonl | cut -c1-900
echo "=== ncu launch list"
export SFVOS_GRAPH=0 SFVOS_LEVEL_STREAMS=0
python bench.py --steps 2 --warmup 3 --no-cpu --no-lib > gpurun_out/r2d_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 2400 --csv --log-file gpurun_out/r2d_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-lib > gpurun_out/r2d_ncu.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/r2d_ncu.log; wc -l gpurun_out/r2d_launches.csv
