#!/bin/bash
# Run the GPU test files in separate processes with timeouts; logs under gpurun_out/.
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout 900 python -m pytest "$@" -q --no-header -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary.txt; tail -12 gpurun_out/$name.log | tee -a gpurun_out/summary.txt; }
for t in "$@"; do run $(basename $t .py) tests/$t.py -m gpu; done
