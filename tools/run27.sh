R=r01
for k in dense_write_kernel interval_sums_kernel da_sca_fwd_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/${R}_$k python bench.py --steps 2 --warmup 3 --no-cpu-baseline --eager-only > /dev/null 2>&1
done
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > gpurun_out/${R}_clocks.csv &
SMI=$!
timeout 900 python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
kill $SMI
tail -c 1200 gpurun_out/${R}_bench.json
ls -la gpurun_out/*.ncu-rep
