#!/bin/bash
# Run GPU test groups in separate processes (a trapped kernel poisons its CUDA context) with timeouts.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout 600 python -m pytest "$@" -q -x --no-header -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary.txt; tail -5 gpurun_out/$name.log | tee -a gpurun_out/summary.txt; }
: > gpurun_out/summary.txt
run simt tests/test_gpu_kernels.py -k "simt"
run misc tests/test_gpu_kernels.py -k "bn_pieces or layout or relu_bwd"
run umma_fprop tests/test_gpu_kernels.py -k "fprop and umma"
run umma_epi tests/test_gpu_kernels.py -k "epilogue and umma"
run umma_dgrad tests/test_gpu_kernels.py -k "dgrad and umma"
run umma_wgrad tests/test_gpu_kernels.py -k "wgrad and umma"
run mod_fp32 tests/test_gpu_slowfast.py -k "fp32"
run mod_bf16 tests/test_gpu_slowfast.py -k "bf16 or roundtrip"
