#!/bin/bash
# Round 2, GPU session F: full suite, bench A/B of the BatchNorm backward from the layer output.
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
echo "=== full suite"; timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/r2f_suite.log 2>&1; echo "exit $?"; tail -8 gpurun_out/r2f_suite.log
cp gpurun_out/parity_report.jsonl gpurun_out/r2f_parity_report.jsonl 2>/dev/null
for v in 1 0; do
echo "=== bench c2 BN_FROM_Y=$v"; SFVOS_BN_FROM_Y=$v timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --no-lib > gpurun_out/r2f_bench_y$v.json 2> gpurun_out/r2f_bench_y$v.err; echo "exit $?"; tail -2 gpurun_out/r2f_bench_y$v.err; cut -c1-330 gpurun_out/r2f_bench_y$v.json
done
