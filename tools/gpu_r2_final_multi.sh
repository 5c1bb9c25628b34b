#!/bin/bash
# Round 2, final multi-GPU check: usage  bash tools/gpu_r2_final_multi.sh N "c2 c5"
mkdir -p gpurun_out
N=${1:-2}
for cfg in ${2:-c2}; do
echo "=== $cfg x$N"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 10 --warmup 3 --config $cfg > gpurun_out/r2z_${cfg}_${N}gpu.json 2> gpurun_out/r2z_${cfg}_${N}gpu.err; echo "exit $?"; grep -v "^\*\|OMP_NUM\|^$\|NCCL version" gpurun_out/r2z_${cfg}_${N}gpu.err | tail -3; grep "^{" gpurun_out/r2z_${cfg}_${N}gpu.json | cut -c1-260
done
