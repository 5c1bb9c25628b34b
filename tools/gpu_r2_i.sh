#!/bin/bash
# Round 2, GPU session I: bench with the streaming e2e, then ncu --set full of the fix-list kernels.
mkdir -p gpurun_out
echo "=== bench c2"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --no-lib > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "exit $?"; tail -2 gpurun_out/r2i_bench.err; cut -c1-300 gpurun_out/r2i_bench.json
echo "=== ncu"
python tools/ncu_small.py > gpurun_out/r2i_plain.log 2>&1 && ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"conv_tstack|wgrad_c32|conv_pair|roi_align" -o gpurun_out/prof_r2i python tools/ncu_small.py > gpurun_out/r2i_ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/r2i_ncu.log; ls -la gpurun_out/prof_r2i.ncu-rep
