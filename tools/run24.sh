timeout 120 python -m pytest tests/test_linear_gpu.py -x -q 2>&1 | tail -4
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:linear_tf32 -s 6 -c 10 python tools/quick_lin.py 2>&1 | grep -E "duration" | awk '{s+=$NF; n++} END {print "case1 cold avg us", s/n}'
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:linear_tf32 -s 6 -c 10 python tools/quick_lin.py 2>&1 | grep -E "duration" | awk '{s+=$NF; n++} END {print "case1 warm avg us", s/n}'
