#!/bin/bash
# Round 2, final single-GPU session: full suite, default bench line, ncu launch list of the final step.
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
echo "=== full suite"; timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/r2q_suite.log 2>&1; echo "exit $?"; tail -4 gpurun_out/r2q_suite.log
cp gpurun_out/parity_report.jsonl gpurun_out/r2q_parity_report.jsonl 2>/dev/null
echo "=== bench default"; timeout 1200 python bench.py > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo "exit $?"; tail -2 gpurun_out/r2q_bench.err; cut -c1-300 gpurun_out/r2q_bench.json
echo "=== pipeline"; timeout 900 python tools/bench_pipeline.py --sequences 4 --frames 24 --sweep-only > gpurun_out/r2q_pipeline.jsonl 2> gpurun_out/r2q_pipeline.err; echo "exit $?"; cut -c1-300 gpurun_out/r2q_pipeline.jsonl
echo "=== ncu launch list"
export SFVOS_GRAPH=0 SFVOS_LEVEL_STREAMS=0
python bench.py --steps 2 --warmup 3 --no-cpu --no-lib > gpurun_out/r2q_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 2400 --csv --log-file gpurun_out/r2q_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-lib > gpurun_out/r2q_ncu.log 2>&1
echo "ncu exit $?"; wc -l gpurun_out/r2q_launches.csv
