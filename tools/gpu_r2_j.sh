#!/bin/bash
# Round 2, GPU session J: kernel + module tests after the epilogue changes, bench C2, conv microbench vs cuDNN.
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
echo "=== tests"; timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/r2j_suite.log 2>&1; echo "exit $?"; tail -6 gpurun_out/r2j_suite.log
cp gpurun_out/parity_report.jsonl gpurun_out/r2j_parity_report.jsonl 2>/dev/null
echo "=== bench c2"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "exit $?"; tail -2 gpurun_out/r2j_bench.err; cut -c1-300 gpurun_out/r2j_bench.json
