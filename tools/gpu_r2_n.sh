#!/bin/bash
# Round 2, 8-GPU check of the final bench (streaming e2e, ncclAvg, fused SGD).
mkdir -p gpurun_out
N=8
echo "=== c2 x8"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2n_c2_8gpu.json 2> gpurun_out/r2n_c2_8gpu.err; echo "exit $?"; grep -v "^\*\|OMP_NUM\|^$\|NCCL version" gpurun_out/r2n_c2_8gpu.err | tail -3; grep "^{" gpurun_out/r2n_c2_8gpu.json | cut -c1-260
echo "=== c2 x1 (same box)"; CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --no-lib > gpurun_out/r2n_c2_1gpu.json 2> gpurun_out/r2n_c2_1gpu.err; echo "exit $?"; grep "^{" gpurun_out/r2n_c2_1gpu.json | cut -c1-260
