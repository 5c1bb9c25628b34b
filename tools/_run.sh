mkdir -p gpurun_out
python tools/ncu_kernels.py > gpurun_out/plain_k.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"conv_pair|conv_umma|conv_tstack|wgrad|roi_align" -c 40 -o gpurun_out/prof_kernels_r1c python tools/ncu_kernels.py > gpurun_out/ncu_k.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_k.log; ls -la gpurun_out/*.ncu-rep
