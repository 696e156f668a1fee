timeout 600 python -m pytest tests/test_forward_gpu.py -x -q 2>&1 | tail -4
for k in 2 0; do
  echo "== kernel=$k"
  FBBEV_POOL_KERNEL=$k timeout 300 python tools/quick_f2.py 2>&1 | grep -E "kernel cold|no flush  |prep" 
done
for cfg in shipped unit_128 fbocc_400; do
FBBEV_POOL_KERNEL=2 timeout 300 python tools/quick_f.py $cfg 1 2>&1 | grep -E "B=1|pool_dense\(kernel only|REF kernel|REF op|algorithmic"
done
