#!/bin/bash
# Round 2, GPU session E: backbone / sequence tests after the cache fix, pipeline phase timing, L2 stream bandwidth probe.
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
echo "=== tests"; timeout 1500 python -m pytest tests/test_gpu_backbone.py tests/test_gpu_sequence.py tests/test_gpu_model.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/r2e_tests.log 2>&1; echo "exit $?"; tail -8 gpurun_out/r2e_tests.log
cp gpurun_out/parity_report.jsonl gpurun_out/r2e_parity_report.jsonl 2>/dev/null
echo "=== pipeline phases"; SFVOS_PIPE_TIMING=1 timeout 900 python tools/bench_pipeline.py --sequences 4 --frames 24 > gpurun_out/r2e_pipeline_phases.jsonl 2> gpurun_out/r2e_pipeline.err; echo "exit $?"; tail -3 gpurun_out/r2e_pipeline.err; cut -c1-1200 gpurun_out/r2e_pipeline_phases.jsonl
echo "=== L2 probe"; timeout 300 python tools/bench_l2.py > gpurun_out/r2e_l2.json 2> gpurun_out/r2e_l2.err; echo "exit $?"; cat gpurun_out/r2e_l2.json; tail -3 gpurun_out/r2e_l2.err
