mkdir -p gpurun_out
SFVOS_GRAPH=0 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/plain_bench.log 2>&1 && SFVOS_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 1700 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
echo "ncu list exit $?"; tail -1 gpurun_out/plain_bench.log | cut -c1-200; wc -l gpurun_out/launches_r1c.csv
