timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ['value','ms_per_step','gpu_launches']}, d['e2e'], d['roofline']['kernel_us'])"
timeout 300 python - <<'PY'
import sys, torch, time
sys.path.insert(0, '.')
import bench
w = bench.Workload('cuda:0', 0)
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); s=torch.cuda.Event(True); e=torch.cuda.Event(True); s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e)/n*1e3
with torch.no_grad():
    print("F total            %.1f us" % t(lambda: w.vt(w.cam, w.feat, w.depth)))
    print("  prepare_from_cams %.1f us" % t(lambda: w.vt.prepare_index_from_cams(*w.cam)))
    print("B total            %.1f us" % t(lambda: w.bp([w.feat], None, lss_bev=w.lss, cam_params=w.cam, pred_img_depth=w.depth)))
    enc = w.bp.transformer.encoder
    print("  point_sampling_fused %.1f us" % t(lambda: enc.point_sampling_fused(w.cam)))
    # CUDA graph of the whole step
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): w.step()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            out = w.step()
    torch.cuda.synchronize()
    print("step eager         %.1f us" % t(lambda: w.step()))
    print("step graph replay  %.1f us" % t(lambda: g.replay()))
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3): w.step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
PY
