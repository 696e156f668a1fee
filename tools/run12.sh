for m in 0 1 2 4 3 6 7; do
echo "== K1 dbg=$m (1=no depth gather, 2=no feat loads, 4=no V store)"
FBBEV_K1_DBG=$m ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"interval_sums" -s 10 -c 3 --csv python tools/quick_f.py fbocc_200 1 2>/dev/null | grep -E "interval_sums" | awk -F'","' '{print $NF}' | tr '\n' ' '
echo
done
