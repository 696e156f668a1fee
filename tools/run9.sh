FBBEV_POOL_TILE=128 ncu --set full --clock-control none --import-source on -k regex:bev_pool_dense_async_kernel -s 8 -c 1 -o gpurun_out/prof_dense_v5_128 python tools/quick_f.py fbocc_200 1 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
