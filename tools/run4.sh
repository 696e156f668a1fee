timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for shp in 128,8 64,4; do echo "== $shp"; FBBEV_POOL_SHAPE=$shp python tools/quick_f2.py; done
