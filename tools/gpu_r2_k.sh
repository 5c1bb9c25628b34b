#!/bin/bash
# experiments on conv_tstack<32,9>: where does the time go?
mkdir -p gpurun_out
for env in "" "SFVOS_TSTACK_DBG=1" "SFVOS_TSTACK_LP=16" "SFVOS_TSTACK_STAGE=2" "SFVOS_TSTACK_G=1"; do
  echo "=== [$env]"; env $env timeout 300 python tools/bench_conv.py --reps 7 fast2 fast3 fast1 fast2+d 2>&1 | tail -4
done > gpurun_out/r2k_tstack_exp.txt 2>&1
cat gpurun_out/r2k_tstack_exp.txt
