"""Where one Gauss-Newton iteration of the cooperative align kernel spends its time (config-2 pair).
usage: python tools/phase_times.py [GICP|POINT_TO_PLANE|POINT_TO_POINT] [iters]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import synthetic  # noqa: E402
import sycl_points_b200 as spx  # noqa: E402

regt = sys.argv[1] if len(sys.argv) > 1 else "GICP"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
q = spx.DeviceQueue(0)
tgt_raw, src_raw, T_gt = synthetic.kitti_pair(42)
vg = spx.VoxelGrid(q, 0.25)
src, tgt = vg.downsampling(spx.PointCloudShared(q, src_raw)), vg.downsampling(spx.PointCloudShared(q, tgt_raw))
ts, tt = spx.KDTree.build(q, src), spx.KDTree.build(q, tgt)
nn_s, nn_t = ts.knn_search(src, 10), tt.knn_search(tgt, 10)
spx.covariance.estimate(nn_s, src)
spx.covariance.estimate(nn_t, tgt)
spx.covariance.estimate_normals(nn_t, tgt)
p = spx.RegistrationParams(reg_type=spx.RegType[regt], max_iterations=iters)
p.robust.type = spx.RobustLossType.HUBER
p.criteria.translation = p.criteria.rotation = 0.0
reg = spx.Registration(q, p)
reg.align(src, tgt, tt)
spx._lib.check(spx.lib().spx_registration_phase_times(reg._h, 1, None, 0))
for _ in range(2):
    reg.align(src, tgt, tt)
buf = np.zeros((iters, 8), np.uint64)
spx._lib.check(spx.lib().spx_registration_phase_times(reg._h, 1, buf.ctypes.data_as(C.c_void_p), iters))
t = buf.astype(np.int64)
names = ["nn fast", "nn coop", "accumulate", "block-reduce", "grid.sync", "fold", "solve"]
d = np.diff(t[:, :8], axis=1) / 1e3
print(f"{regt}: {src.size()} source points, loop {reg.last_timing()['loop_ms'] * 1e3 / iters:.1f} us/iter")
print("phase (us, latest block):  " + "  ".join(f"{n:>12s}" for n in names) + "   total")
for it in range(iters):
    print(f"  iter {it:2d}                  " + "  ".join(f"{v:12.1f}" for v in d[it]) + f"   {d[it].sum():6.1f}")
print("  mean                     " + "  ".join(f"{v:12.1f}" for v in d[1:].mean(0)) + f"   {d[1:].sum(1).mean():6.1f}")


# work-list statistics of the batched kernel: {list size, list queries that end without a neighbour inside max_corr}
full = np.zeros(64 * 8 + 8192 * 5, np.uint64)
spx._lib.check(spx.lib().spx_registration_phase_times(reg._h, 1, full.ctypes.data_as(C.c_void_p), -1))
w = full[64 * 8:64 * 8 + 2 * iters].reshape(-1, 2).astype(np.int64)
for it in range(iters):
    print(f"  iter {it:2d}: work list {w[it, 0]:7d} queries ({100.0 * w[it, 0] / src.size():5.1f} % of the source), "
          f"{w[it, 1]:7d} of them end beyond max_corr")
