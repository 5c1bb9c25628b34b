#!/bin/bash
# Round 2, session r: coalescing epilogue (common.cuh epi_block) in conv_umma / conv_pair - parity subset + A/B timing.
mkdir -p gpurun_out
echo "=== kernel tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_roi_mask.py tests/test_gpu_box_head.py tests/test_gpu_fullsize.py -m gpu -x -q --no-header -p no:cacheprovider > gpurun_out/r2r_tests.log 2>&1; echo "exit $?"; tail -15 gpurun_out/r2r_tests.log
for s in 0 1; do
  echo "=== SFVOS_EPI_STAGE=$s"
  SFVOS_EPI_STAGE=$s timeout 600 python tools/bench_conv.py convt4 maskconv maskconv+d fc6 slow1 slow1+d slow3 slow3+d f2s1 f2s2 2>&1 | tee -a gpurun_out/r2r_ab.txt
done
