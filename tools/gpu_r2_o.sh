#!/bin/bash
mkdir -p gpurun_out
N=8
for sp in 1 0; do
echo "=== c2 x8 SFVOS_DP_SPLIT=$sp"; SFVOS_DP_SPLIT=$sp timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540+sp)) bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2o_c2_8gpu_split$sp.json 2> gpurun_out/r2o_c2_8gpu_split$sp.err; echo "exit $?"; grep "^{" gpurun_out/r2o_c2_8gpu_split$sp.json | cut -c1-220
done
