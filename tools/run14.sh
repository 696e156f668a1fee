timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -3
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
    print("  get_lidar_coor   %.1f us" % t(lambda: w.vt.get_lidar_coor(*w.cam)))
    coor = w.vt.get_lidar_coor(*w.cam)
    print("  prepare_index    %.1f us" % t(lambda: w.vt.prepare_index(coor)))
    idx = w.vt.prepare_index(coor)
    print("  pool (plan+pool+permute) %.1f us" % t(lambda: w.vt._pool(idx, w.depth, w.feat)))
    print("  inverse(3x3 x6)  %.1f us" % t(lambda: torch.inverse(w.cam[3])))
    print("B total            %.1f us" % t(lambda: w.bp([w.feat], None, lss_bev=w.lss, cam_params=w.cam, pred_img_depth=w.depth)))
    enc = w.bp.transformer.encoder
    ref3d = enc.get_reference_points(200, 200, dim='3d', device='cuda:0')
    print("  get_reference_points %.1f us" % t(lambda: enc.get_reference_points(200, 200, dim='3d', device='cuda:0')))
    print("  point_sampling   %.1f us" % t(lambda: enc.point_sampling(ref3d, enc.pc_range, None, cam_params=w.cam)))
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3): w.bp([w.feat], None, lss_bev=w.lss, cam_params=w.cam, pred_img_depth=w.depth)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
PY
