# Produces the round's GPU evidence (run under gpurun, 1 GPU): parity suite,
# smoke, the bench line, the ncu launch list of the eager step and one
# `--set full` capture per hand-written hot kernel.  Summaries for profiles/
# are made afterwards with tools/summarize_round.sh.
R=${1:-r02}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --durations=8 2>&1 | tail -16
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > gpurun_out/${R}_clocks.csv &
SMI=$!
timeout 900 python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
kill $SMI
tail -c 1500 gpurun_out/${R}_bench.json
# every launch of ~3 eager steps (warm-up steps skipped)
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 200 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --eager-only > gpurun_out/${R}_bench_under_ncu.log 2>&1
for k in dense_write_kernel interval_sums_kernel linear_tf32_kernel ffn_tf32_kernel da_sca_smem_kernel msda_fused_fwd_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o gpurun_out/${R}_$k python bench.py --steps 2 --warmup 3 --no-cpu-baseline --eager-only > /dev/null 2>&1
done
if [ "${SKIP_HISTORY:-0}" != "1" ]; then
ncu --set full --clock-control none --import-source on -k regex:history_warp_kernel -s 3 -c 1 -f -o gpurun_out/${R}_history_warp_kernel python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-frames16 > /dev/null 2>&1
fi
# the other BASELINE.json configs as bench lines
for c in unit frames16 bwd_only large; do
  timeout 600 python bench.py --config $c > gpurun_out/${R}_bench_$c.json 2> gpurun_out/${R}_bench_$c.err
done
ls -la gpurun_out/ | tail -14
