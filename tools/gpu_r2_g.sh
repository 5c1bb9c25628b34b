#!/bin/bash
# Round 2, GPU session G: native ResNet body tests, model-level tests, C4 pipeline with phases.
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
echo "=== tests"; timeout 1500 python -m pytest tests/test_gpu_backbone.py tests/test_gpu_sequence.py tests/test_gpu_model.py tests/test_gpu_callers.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/r2g_tests.log 2>&1; echo "exit $?"; tail -12 gpurun_out/r2g_tests.log
cp gpurun_out/parity_report.jsonl gpurun_out/r2g_parity_report.jsonl 2>/dev/null
grep resnet_body gpurun_out/r2g_parity_report.jsonl
echo "=== pipeline phases (native body + FPN + RPN head)"; SFVOS_PIPE_TIMING=1 timeout 900 python tools/bench_pipeline.py --sequences 4 --frames 24 > gpurun_out/r2g_pipeline_phases.jsonl 2> gpurun_out/r2g_pipeline.err; echo "exit $?"; tail -3 gpurun_out/r2g_pipeline.err; cut -c1-700 gpurun_out/r2g_pipeline_phases.jsonl
echo "=== pipeline (no phase sync)"; timeout 900 python tools/bench_pipeline.py --sequences 4 --frames 24 > gpurun_out/r2g_pipeline.jsonl 2>> gpurun_out/r2g_pipeline.err; echo "exit $?"; cut -c1-400 gpurun_out/r2g_pipeline.jsonl
