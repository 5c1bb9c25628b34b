#!/bin/bash
# Round 2, GPU session M: default bench line (N=1, single graph), reference arm, smoke, 20x full GPU suite.
mkdir -p gpurun_out
echo "=== bench default"; timeout 1200 python bench.py > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "exit $?"; tail -2 gpurun_out/r2m_bench.err; cut -c1-300 gpurun_out/r2m_bench.json
echo "=== bench reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2m_bench_ref.json 2> gpurun_out/r2m_bench_ref.err; echo "exit $?"; cut -c1-300 gpurun_out/r2m_bench_ref.json
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2m_smoke.log 2>&1; echo "exit $?"; tail -3 gpurun_out/r2m_smoke.log
echo "=== 20x suite"; : > gpurun_out/r2m_loop20.log
for i in $(seq 1 20); do timeout 900 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider 2>&1 | tail -1 | sed "s/^/run $i: /" >> gpurun_out/r2m_loop20.log; done
cat gpurun_out/r2m_loop20.log
