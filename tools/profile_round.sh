# Produces the ncu evidence for profiles/ (run under gpurun, 1 GPU).
set -x
R=${1:-r01}
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/${R}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dense_write_kernel -s 4 -c 1 -f -o gpurun_out/${R}_dense_write python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:interval_sums_kernel -s 4 -c 1 -f -o gpurun_out/${R}_interval_sums python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:da_sca_fwd_kernel -s 4 -c 1 -f -o gpurun_out/${R}_da_sca python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > /dev/null 2>&1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > gpurun_out/${R}_clocks.csv &
SMI=$!
python bench.py --steps 30 --warmup 5 > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
kill $SMI
tail -c 3000 gpurun_out/${R}_bench.json
