timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench.err | tail -c 2500
FBBEV_TORCH_LINEAR=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('torch-linear: ms_per_step', d['ms_per_step'], 'value', d['value'])"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:linear_tf32 -c 12 python tools/quick_lin.py 2>&1 | grep -E "linear_tf32_kernel|duration" | paste - - | awk '{print $2,$3, $(NF)}' | sort | uniq -c
