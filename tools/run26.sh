for t in 128 64; do for cfg in shipped unit_128 fbocc_400; do
echo "== T=$t $cfg"; FBBEV_POOL_TILE=$t timeout 200 python tools/quick_f.py $cfg 1 2>&1 | grep -E "pool_dense\(kernel only"
done; done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | tail -c 1800
tail -3 gpurun_out/bench.err
