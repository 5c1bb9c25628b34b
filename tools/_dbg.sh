( time python bench.py --steps 20 --warmup 5 > gpurun_out/r2z2_bench.json 2> gpurun_out/r2z2_bench.err ) 2>&1 | tail -4; cut -c1-200 gpurun_out/r2z2_bench.json
( time python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2z2_ref.json 2> gpurun_out/r2z2_ref.err ) 2>&1 | tail -4; cut -c1-300 gpurun_out/r2z2_ref.json
