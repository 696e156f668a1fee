set -x
ncu --set full --clock-control none --import-source on -k regex:bev_pool_dense_kernel -s 8 -c 2 -o gpurun_out/prof_dense_r1a python tools/quick_f.py fbocc_200 1 > gpurun_out/ncu_dense.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 120 --csv --log-file gpurun_out/launches_quick_f.csv python tools/quick_f.py fbocc_200 1 > /dev/null 2>&1
tail -5 gpurun_out/ncu_dense.log
