#!/bin/bash
mkdir -p gpurun_out
for env in "SFVOS_PAIR_ASTAGES=2 SFVOS_PAIR_BSTAGES=3" "SFVOS_PAIR_ASTAGES=2 SFVOS_PAIR_BSTAGES=4" "SFVOS_PAIR_ASTAGES=2 SFVOS_PAIR_BSTAGES=5" "SFVOS_PAIR_ASTAGES=2 SFVOS_PAIR_BSTAGES=6" "SFVOS_PAIR_ASTAGES=2 SFVOS_PAIR_BSTAGES=8" "SFVOS_PAIR_ASTAGES=3 SFVOS_PAIR_BSTAGES=6" "SFVOS_PAIR_ASTAGES=3 SFVOS_PAIR_BSTAGES=4"; do
  echo "=== [$env]"; env $env timeout 300 python tools/bench_conv.py --reps 9 slow1 slow2+d slow3 2>&1 | tail -3 | cut -c1-70
done > gpurun_out/r2p_pair_stages2.txt 2>&1
cat gpurun_out/r2p_pair_stages2.txt
