FBBEV_POOL_STREAM=1 timeout 200 python -m pytest tests/test_forward_gpu.py -x -q 2>&1 | tail -2
for s in 0 1; do echo "== STREAM=$s"; FBBEV_POOL_STREAM=$s timeout 100 python tools/quick_f2.py 2>&1 | grep -E "kernel cold|kernel back"; done
