FBBEV_POOL_SHAPE=128,8 ncu --set full --clock-control none --import-source on -k regex:bev_pool_dense_kernel -s 8 -c 1 -o gpurun_out/prof_dense_v2_128x8 python tools/quick_f.py fbocc_200 1 > /dev/null 2>&1
FBBEV_POOL_SHAPE=64,4 ncu --set full --clock-control none --import-source on -k regex:bev_pool_dense_kernel -s 8 -c 1 -o gpurun_out/prof_dense_v2_64x4 python tools/quick_f.py fbocc_200 1 > /dev/null 2>&1
ls -la gpurun_out/
