timeout 200 python tools/bench_conv.py --reps 9 fast2 fast3 fast2+d fast3+d f2s1+d f2s2+d f2s1 f2s2 convt 2>&1 | tail -9
