set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for shp in 128,8 128,4 64,8 64,4 32,8 32,4; do
  echo "=== SHAPE $shp"
  FBBEV_POOL_SHAPE=$shp timeout 300 python tools/quick_f.py fbocc_200 1 2>&1 | grep -E "kernel only|algorithmic|plan\+kernel"
done
for shp in 128,8 64,8 32,4; do
  echo "=== SHAPE $shp shipped / unit"
  FBBEV_POOL_SHAPE=$shp timeout 300 python tools/quick_f.py shipped 1 2>&1 | grep -E "pool_dense\(kernel only|REF kernel"
  FBBEV_POOL_SHAPE=$shp timeout 300 python tools/quick_f.py unit_128 1 2>&1 | grep -E "pool_dense\(kernel only|REF kernel"
done
