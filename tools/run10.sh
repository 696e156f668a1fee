timeout 300 python -m pytest tests/test_forward_gpu.py -x -q 2>&1 | tail -3
echo "== split tile=128"
timeout 300 python tools/quick_f2.py 2>&1 | grep -E "kernel cold|no flush  |prep" 
for cfg in shipped unit_128 fbocc_400; do
timeout 300 python tools/quick_f.py $cfg 1 2>&1 | grep -E "B=1|pool_dense\(kernel only|REF kernel|REF op|algorithmic"
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pool_split|split_plan" -s 20 -c 6 --csv python tools/quick_f.py fbocc_200 1 2>/dev/null | grep -E "pool_split|split_plan" | awk -F'","' '{print $5, $NF}' | head -12
