#!/bin/bash
# Round 2, final single-GPU evidence: full suite, default bench line (+ reference arm), C4 pipeline, L2 probe, ncu launch list of
# the final step, ncu --set full of the hot kernels.  Outputs: gpurun_out/r2z_*.
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
echo "=== full suite"; timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/r2z_suite.log 2>&1; echo "exit $?"; tail -3 gpurun_out/r2z_suite.log
cp gpurun_out/parity_report.jsonl gpurun_out/r2z_parity_report.jsonl 2>/dev/null
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2z_smoke.log 2>&1; echo "exit $?"; tail -2 gpurun_out/r2z_smoke.log
echo "=== bench default"; timeout 1500 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "exit $?"; tail -2 gpurun_out/r2z_bench.err; cut -c1-300 gpurun_out/r2z_bench.json
echo "=== bench reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2z_bench_ref.json 2> gpurun_out/r2z_bench_ref.err; echo "exit $?"; cut -c1-300 gpurun_out/r2z_bench_ref.json
echo "=== pipeline"; timeout 900 python tools/bench_pipeline.py --sequences 4 --frames 24 --sweep-only > gpurun_out/r2z_pipeline.jsonl 2> gpurun_out/r2z_pipeline.err; echo "exit $?"; cut -c1-300 gpurun_out/r2z_pipeline.jsonl
echo "=== l2 probe"; timeout 300 python tools/bench_l2.py > gpurun_out/r2z_l2.json 2>&1; cat gpurun_out/r2z_l2.json
echo "=== timeline"; timeout 600 python tools/profile_timeline.py --out gpurun_out/r2z_timeline.csv > gpurun_out/r2z_timeline.txt 2>&1; head -3 gpurun_out/r2z_timeline.txt
echo "=== ncu launch list"
( export SFVOS_GRAPH=0 SFVOS_LEVEL_STREAMS=0
python bench.py --steps 2 --warmup 3 --no-cpu --no-lib > gpurun_out/r2z_plain.log 2>&1 && timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 2400 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-lib > gpurun_out/r2z_ncu.log 2>&1
echo "ncu exit $?"; wc -l gpurun_out/r2z_launches.csv )
echo "=== ncu full: hot kernels"
python tools/ncu_small.py > gpurun_out/r2z_small_plain.log 2>&1 && timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"conv_tstack|wgrad_c32|conv_pair|roi_align" -o gpurun_out/prof_r2z_small python tools/ncu_small.py > gpurun_out/r2z_small_ncu.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/prof_r2z_small.ncu-rep
python tools/ncu_step.py > gpurun_out/r2z_step_plain.log 2>&1 && timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"nchw_to_nhwc_bf16|mask_logits_relu_bwd_c256|conv_umma_kernel<64>" -c 8 -o gpurun_out/prof_r2z_step python tools/ncu_step.py > gpurun_out/r2z_step_ncu.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/prof_r2z_step.ncu-rep
