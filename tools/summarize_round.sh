# After tools/profile_round.sh: turn gpurun_out/ captures into the tracked summaries under profiles/.
R=${1:-r02}
python tools/summarize_ncu.py launches gpurun_out/${R}_launches.csv profiles/${R}_launches.md
for k in dense_write_kernel interval_sums_kernel linear_tf32_kernel ffn_tf32_kernel da_sca_smem_kernel msda_fused_fwd_kernel history_warp_kernel; do
  [ -f gpurun_out/${R}_$k.ncu-rep ] && python tools/summarize_ncu.py kernel gpurun_out/${R}_$k.ncu-rep profiles/${R}_$k.json
done
python - <<PY
import json
r = "${R}"
try:
    a = json.load(open(f"profiles/{r}_dense_write_kernel.json"))
    b = json.load(open(f"profiles/{r}_interval_sums_kernel.json"))
    json.dump({"kernel": "interval_sums_kernel + dense_write_kernel (one dense pooling call)",
               "source": f"profiles/{r}_interval_sums_kernel.json + profiles/{r}_dense_write_kernel.json "
                         "(ncu --set full, one launch each, config 2, caches flushed before each launch)",
               "dram_bytes_per_launch": a["dram_bytes_per_launch"] + b["dram_bytes_per_launch"]},
              open("profiles/pool_dense_traffic.json", "w"), indent=1)
except Exception as e:
    print("traffic summary skipped:", e)
PY
cp gpurun_out/${R}_bench.json profiles/${R}_bench_line.json 2>/dev/null
for c in unit frames16 bwd_only large; do [ -s gpurun_out/${R}_bench_$c.json ] && cp gpurun_out/${R}_bench_$c.json profiles/${R}_bench_$c.json; done
cp gpurun_out/${R}_clocks.csv profiles/${R}_clocks.csv 2>/dev/null
ls -la profiles/
