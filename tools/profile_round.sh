# Produces the round's GPU evidence (run under gpurun, 1 GPU): parity suite,
# smoke, the bench line, the ncu launch list of the eager step and one
# `--set full` capture per hand-written hot kernel.  Summaries for profiles/
# are made afterwards with tools/summarize_ncu.py.
R=${1:-r01}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu --durations=8 2>&1 | tail -16
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > gpurun_out/${R}_clocks.csv &
SMI=$!
timeout 900 python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
kill $SMI
tail -c 2600 gpurun_out/${R}_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 360 -c 260 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --eager-only > gpurun_out/${R}_bench_under_ncu.log 2>&1
for k in dense_write_kernel interval_sums_kernel linear_tf32_kernel da_sca_fwd_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o gpurun_out/${R}_$k python bench.py --steps 2 --warmup 3 --no-cpu-baseline --eager-only > /dev/null 2>&1
done
ls -la gpurun_out/ | tail -12
