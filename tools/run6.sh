timeout 600 python -m pytest tests/test_forward_gpu.py -x -q 2>&1 | tail -4
for k in 0 1; do for shp in 128,8 64,4 64,8; do
  echo "== kernel=$k shape=$shp"
  FBBEV_POOL_KERNEL=$k FBBEV_POOL_SHAPE=$shp timeout 300 python tools/quick_f2.py 2>&1 | grep -E "kernel cold|no flush  " | head -3
done; done
FBBEV_POOL_KERNEL=1 FBBEV_POOL_SHAPE=64,4 timeout 300 python tools/quick_f.py shipped 1 2>&1 | grep -E "pool_dense\(kernel only|REF kernel"
FBBEV_POOL_KERNEL=1 FBBEV_POOL_SHAPE=64,4 timeout 300 python tools/quick_f.py unit_128 1 2>&1 | grep -E "pool_dense\(kernel only|REF kernel"
