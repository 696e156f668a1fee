timeout 600 python -m pytest tests/test_forward_gpu.py -x -q 2>&1 | tail -4
for t in 64 128 32; do
  echo "== async tile=$t"
  FBBEV_POOL_TILE=$t timeout 300 python tools/quick_f2.py 2>&1 | grep -E "kernel cold|no flush  |prep" 
done
echo "== generic kernel (FBBEV_POOL_KERNEL=0)"
FBBEV_POOL_KERNEL=0 timeout 300 python tools/quick_f2.py 2>&1 | grep -E "kernel cold|kernel back"
for cfg in shipped unit_128 fbocc_400; do
timeout 300 python tools/quick_f.py $cfg 1 2>&1 | grep -E "B=1|pool_dense\(kernel only|REF kernel|REF op|algorithmic"
done
