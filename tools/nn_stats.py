"""Per-query work of the k = 1 index search on the config-2 pair (tuning aid)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import synthetic  # noqa: E402
import sycl_points_b200 as spx  # noqa: E402

q = spx.DeviceQueue(0)
tgt_raw, src_raw, T_gt = synthetic.kitti_pair(42)
vg = spx.VoxelGrid(q, 0.25)
src, tgt = vg.downsampling(spx.PointCloudShared(q, src_raw)), vg.downsampling(spx.PointCloudShared(q, tgt_raw))
cell = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
tt = spx.KDTree.build(q, tgt, cell_size=cell)
print(tt.info())
for name, radius, qs in (("bounded 2.0 src->tgt", 2.0, src), ("unbounded src->tgt", 0.0, src), ("unbounded self", 0.0, tgt)):
    st = spx.DeviceArray(q, (qs.size(), 4), np.uint32)
    spx._lib.check(spx.lib().spx_index_nn_stats(tt.handle, qs.points.ptr, qs.size(), None, radius, st.ptr))
    h = st.download()
    print("==", name)
    for col, nm in enumerate(("segments", "candidates", "shells", "last_level")):
        v = h[:, col]
        print(f"  {nm:11s} mean {v.mean():8.1f}  p50 {np.percentile(v,50):6.0f} p90 {np.percentile(v,90):6.0f} "
              f"p99 {np.percentile(v,99):7.0f} p99.9 {np.percentile(v,99.9):8.0f} max {v.max():8d}")
    w = h.reshape(-1)[: (len(h) // 32) * 32 * 4].reshape(-1, 32, 4)
    print("  per-warp max candidates: mean %.0f p99 %.0f ; sum of warp-max / sum = %.2f" %
          (w[:, :, 1].max(1).mean(), np.percentile(w[:, :, 1].max(1), 99), w[:, :, 1].max(1).sum() * 32 / max(h[:, 1].sum(), 1)))
    print("  level histogram:", np.bincount(h[:, 3], minlength=6))
