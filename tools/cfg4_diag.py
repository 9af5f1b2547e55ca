"""Config 4 (dense pair) diagnostics: correspondence-distance distribution after one iteration and the align's time
per iteration for a few iteration counts.  usage: python tools/cfg4_diag.py [iters,iters,...]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import synthetic  # noqa: E402
import sycl_points_b200 as spx  # noqa: E402

q = spx.DeviceQueue(0)
tgt_raw, src_raw, T_gt = synthetic.dense_pair(42)
vg = spx.VoxelGrid(q, 0.05)
src_full = vg.downsampling(spx.PointCloudShared(q, src_raw)).points_host()[:2_000_000]
tgt_full = vg.downsampling(spx.PointCloudShared(q, tgt_raw)).points_host()[:2_000_000]
del src_raw, tgt_raw
src, tgt = spx.PointCloudShared(q, src_full), spx.PointCloudShared(q, tgt_full)
ts, tt = spx.KDTree.build(q, src), spx.KDTree.build(q, tgt)
print("target index", tt.info())
spx.covariance.estimate(ts.knn_search(src, 10), src)
spx.covariance.estimate(tt.knn_search(tgt, 10), tgt)
ITERS = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1, 2, 3, 5, 10, 20, 40]
for iters in ITERS:
    p = spx.RegistrationParams(reg_type=spx.RegType.GICP, max_iterations=iters)
    p.robust.type = spx.RobustLossType.HUBER
    p.criteria.translation = p.criteria.rotation = 0.0
    reg = spx.Registration(q, p)
    reg.align(src, tgt, tt)
    reg.align(src, tgt, tt)
    t = reg.last_timing()
    print(f"{iters} iterations: loop {t['loop_ms']:.3f} ms -> {t['loop_ms'] / iters:.3f} ms/iter; "
          f"kept without a search {reg.kept_correspondences()} of {iters * src.size()} correspondences")
    idx_p, dist_p, n = C.c_void_p(), C.c_void_p(), C.c_size_t()
    spx._lib.check(spx.lib().spx_registration_neighbors(reg._h, C.byref(idx_p), C.byref(dist_p), C.byref(n)))
    d = np.empty(n.value, np.float32)
    spx._lib.check(spx.lib().spx_memcpy_d2h(q.handle, d.ctypes.data_as(C.c_void_p), dist_p, d.nbytes))
    q.wait()
    r = np.sqrt(np.minimum(d, 1e6))
    print(f"  correspondence distance (last search): beyond 2 m {np.sum(d > 4.0)} of {len(d)}; "
          f"p50 {np.percentile(r, 50):.3f} p90 {np.percentile(r, 90):.3f} p99 {np.percentile(r, 99):.3f} "
          f"p99.9 {np.percentile(r, 99.9):.3f} max {r.max():.3f};  > 0.2 m: {np.sum(r > 0.2)}, > 0.5 m: {np.sum(r > 0.5)}, > 1 m: {np.sum(r > 1.0)}")
