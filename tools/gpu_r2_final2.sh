#!/bin/bash
# Round 2, last refresh of the evidence with the final code: suite, driver-style bench line (K = 20), reference arm, ncu launch list.
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
echo "=== full suite"; timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/r2zz_suite.log 2>&1; echo "exit $?"; tail -2 gpurun_out/r2zz_suite.log
cp gpurun_out/parity_report.jsonl gpurun_out/r2zz_parity_report.jsonl 2>/dev/null
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2zz_smoke.log 2>&1; echo "exit $?"; tail -1 gpurun_out/r2zz_smoke.log
echo "=== bench"; timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2zz_bench.json 2> gpurun_out/r2zz_bench.err; echo "exit $?"; cut -c1-220 gpurun_out/r2zz_bench.json
echo "=== bench reference arm"; timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2zz_bench_ref.json 2> gpurun_out/r2zz_bench_ref.err; echo "exit $?"; cut -c1-200 gpurun_out/r2zz_bench_ref.json
echo "=== timeline"; timeout 600 python tools/profile_timeline.py --out gpurun_out/r2zz_timeline.csv > gpurun_out/r2zz_timeline.txt 2>&1; grep "kernel time" gpurun_out/r2zz_timeline.txt
echo "=== ncu launch list"
( export SFVOS_GRAPH=0 SFVOS_LEVEL_STREAMS=0
python bench.py --steps 2 --warmup 3 --no-cpu --no-lib > gpurun_out/r2zz_plain.log 2>&1 && timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 2400 --csv --log-file gpurun_out/r2zz_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-lib > gpurun_out/r2zz_ncu.log 2>&1
echo "ncu exit $?"; wc -l gpurun_out/r2zz_launches.csv )
