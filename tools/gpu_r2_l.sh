#!/bin/bash
# Round 2, 4-GPU session: C2 / C5 scaling points, plus C3 on one GPU.
mkdir -p gpurun_out
run() { name=$1; n=$2; shift 2; echo "=== $name"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520 + RANDOM % 200)) "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "exit $?"; grep -v "^\*\|OMP_NUM\|^$\|NCCL version" gpurun_out/$name.err | tail -3; grep "^{" gpurun_out/$name.json | cut -c1-260; }
run r2l_c2_4gpu 4 bench.py --gpus 4 --steps 10 --warmup 3
run r2l_c5_4gpu 4 bench.py --gpus 4 --steps 5 --warmup 3 --config c5
echo "=== c3 b1"; CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu --no-lib > gpurun_out/r2l_c3_b1.json 2> gpurun_out/r2l_c3_b1.err; echo "exit $?"; tail -2 gpurun_out/r2l_c3_b1.err; cut -c1-260 gpurun_out/r2l_c3_b1.json
echo "=== c3 b4"; CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py --config c3 --clips 4 --micro 4 --steps 5 --warmup 3 --no-cpu --no-lib > gpurun_out/r2l_c3_b4.json 2> gpurun_out/r2l_c3_b4.err; echo "exit $?"; tail -2 gpurun_out/r2l_c3_b4.err; cut -c1-260 gpurun_out/r2l_c3_b4.json
