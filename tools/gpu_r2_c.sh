#!/bin/bash
# Round 2, GPU session C (2 GPUs): full suite on GPU 0 (incl. the two-device test), bench C2 at 1 and 2 GPUs, C5 at 2 GPUs.
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
echo "=== full suite"; timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/r2c_suite.log 2>&1; echo "exit $?"; tail -15 gpurun_out/r2c_suite.log
cp gpurun_out/parity_report.jsonl gpurun_out/r2c_parity_report.jsonl 2>/dev/null
echo "=== bench c2 x1"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --no-lib > gpurun_out/r2c_bench_c2_1gpu.json 2> gpurun_out/r2c_bench_c2_1gpu.err; echo "exit $?"; tail -3 gpurun_out/r2c_bench_c2_1gpu.err; cut -c1-400 gpurun_out/r2c_bench_c2_1gpu.json
echo "=== bench c2 x2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2c_bench_c2_2gpu.json 2> gpurun_out/r2c_bench_c2_2gpu.err; echo "exit $?"; tail -3 gpurun_out/r2c_bench_c2_2gpu.err; cut -c1-400 gpurun_out/r2c_bench_c2_2gpu.json
echo "=== bench c5 x2"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --config c5 > gpurun_out/r2c_bench_c5_2gpu.json 2> gpurun_out/r2c_bench_c5_2gpu.err; echo "exit $?"; tail -3 gpurun_out/r2c_bench_c5_2gpu.err; cut -c1-400 gpurun_out/r2c_bench_c5_2gpu.json
