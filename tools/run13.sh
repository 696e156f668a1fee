ncu --set full --clock-control none --import-source on -k regex:interval_sums_kernel -s 40 -c 1 -o gpurun_out/prof_isums3 python tools/quick_f.py fbocc_200 1 > /dev/null 2>&1
