set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -30
timeout 300 python tools/quick_f.py fbocc_200 1 2>&1 | tail -20
timeout 300 python tools/quick_f.py shipped 1 2>&1 | tail -20
timeout 300 python tools/quick_f.py unit_128 1 2>&1 | tail -20
