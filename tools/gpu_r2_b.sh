#!/bin/bash
# Round 2, GPU session B: full suite (+ measured report), 20x repeat, bench (default) + library baselines, launch list.
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
echo "=== full suite" ; timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/r2b_suite.log 2>&1; echo "exit $?"; tail -30 gpurun_out/r2b_suite.log
cp gpurun_out/parity_report.jsonl gpurun_out/r2b_parity_report.jsonl 2>/dev/null
echo "=== bench"; timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "exit $?"; tail -5 gpurun_out/r2b_bench.err; cut -c1-1500 gpurun_out/r2b_bench.json
echo "=== 20x suite"; : > gpurun_out/r2b_loop20.log
for i in $(seq 1 20); do timeout 900 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider 2>&1 | tail -1 | sed "s/^/run $i: /" >> gpurun_out/r2b_loop20.log; done
cat gpurun_out/r2b_loop20.log
