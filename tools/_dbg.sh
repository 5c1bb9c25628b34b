for ad in f32 bf16; do BENCH_DX_BF16=1 BENCH_ADDEND=$ad timeout 300 python tools/bench_conv.py f2s1+d f2s2+d 2>&1 | sed "s/$/ addend=$ad/"; done
timeout 300 python tools/bench_conv.py fast1 fast2 fast3 fast2+d f2s1+d 2>&1
timeout 600 python tools/profile_timeline.py --out gpurun_out/r2x_timeline.csv 2>&1 | grep "kernel time"
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_slowfast.py -m gpu -x -q --no-header -p no:cacheprovider 2>&1 | tail -2
