N="ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg --clock-control none -k regex:linear_tf32 -s 3 -c 3"
echo base; $N python tools/quick_lin1.py 2>&1 | grep -E "duration|cycles_active"
for v in LIN_NOLOAD LIN_NOEPI; do echo $v; FBBEV_LIB=$PWD/build/var_$v/libfbbev_b200.so $N python tools/quick_lin1.py 2>&1 | grep -E "duration|cycles_active"; done
