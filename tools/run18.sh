ncu --set full --clock-control none --import-source on -k regex:interval_sums_kernel -s 8 -c 1 -f -o gpurun_out/k1 python tools/quick_f2.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:dense_write_kernel -s 8 -c 1 -f -o gpurun_out/k2 python tools/quick_f2.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
