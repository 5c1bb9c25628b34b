#!/bin/bash
# N consecutive runs of the GPU suite on one box (determinism / flakiness check): bash tools/gpu_r2_loop.sh 10
mkdir -p gpurun_out
: > gpurun_out/r2z_loop.log
for i in $(seq 1 ${1:-10}); do
  r=$(timeout 600 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider 2>&1 | tail -1)
  echo "run $i: $r" | tee -a gpurun_out/r2z_loop.log
done
