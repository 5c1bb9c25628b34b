#!/bin/bash
# Round 2, multi-GPU session (N = $1): C2 and C5 through bench.py under torchrun, C4 pipeline sharded by sequence.
N=${1:-8}
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + RANDOM % 200)) "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "exit $?"; grep -v "^\*\|OMP_NUM\|^$\|NCCL version" gpurun_out/$name.err | tail -3; grep "^{" gpurun_out/$name.json | cut -c1-260; }
run r2h_c2_${N}gpu bench.py --gpus $N --steps 10 --warmup 3
run r2h_c5_${N}gpu bench.py --gpus $N --steps 5 --warmup 3 --config c5
run r2h_c4_${N}gpu tools/bench_pipeline.py --sequences 16 --frames 60 --sweep-only
