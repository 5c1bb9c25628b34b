mkdir -p gpurun_out
python bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 exit $?"; cut -c1-700 gpurun_out/bench_2gpu.json; tail -3 gpurun_out/bench_2gpu.err
python tools/profile_slowfast.py --sp 4 --fp 32 --B 1 --steps 3 --no-prof 2>&1 | tail -1
python tools/profile_slowfast.py --sp 4 --fp 32 --B 2 --steps 3 --no-prof 2>&1 | tail -1
python tools/profile_slowfast.py --sp 2 --fp 16 --B 8 --steps 3 --no-prof 2>&1 | tail -1
python tools/profile_slowfast.py --sp 1 --fp 8 --B 8 --steps 3 --no-prof 2>&1 | tail -1
