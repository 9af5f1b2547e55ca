"""Diagnostic: where do the voxel map's covariances differ from the sequential oracle?  (GPU box)"""
import ctypes as C
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle, synthetic
import sycl_points_b200 as spx
from sycl_points_b200 import _lib

q = spx.DeviceQueue(0)
tgt_raw, _, _ = synthetic.kitti_pair(100, sweeps=1, azimuth_steps=1024)
cloud = spx.VoxelGrid(q, 0.25).downsampling(spx.PointCloudShared(q, tgt_raw))
tree = spx.KDTree.build(q, cloud)
spx.covariance.estimate(tree.knn_search(cloud, 10), cloud)
n = cloud.size()
c16 = cloud.covs.download(n)
pts = cloud.points_host()
# 1. log / exp in isolation
m9 = np.ascontiguousarray(c16.reshape(n, 4, 4)[:, :3, :3]).reshape(n, 9)  # col-major 3x3 (symmetric anyway)
d_in, d_out = spx.DeviceArray.from_host(q, m9), spx.DeviceArray(q, (n, 9), np.float32)
_lib.check(_lib.lib().spx_spd_function(q.handle, d_in.ptr, n, 1, 1e-6, d_out.ptr))
glog = d_out.download()
olog = np.array([oracle.spd_function(m9[i].reshape(3, 3).T, True).T.reshape(9) for i in range(n)])
e = np.abs(glog - olog).max(1)
print("log_spd device vs oracle: max abs", e.max(), "n>1e-6:", int((e > 1e-6).sum()), "of", n)
for i in np.argsort(e)[-3:]:
    print(" worst", i, e[i], "\n", m9[i].reshape(3, 3), "\n gpu", glog[i].reshape(3, 3), "\n orc", olog[i].reshape(3, 3))
d_in2 = spx.DeviceArray.from_host(q, olog.astype(np.float32))
_lib.check(_lib.lib().spx_spd_function(q.handle, d_in2.ptr, n, 0, 0.0, d_out.ptr))
gexp = d_out.download()
oexp = np.array([oracle.spd_function(olog[i].reshape(3, 3).T, False).T.reshape(9) for i in range(n)])
e2 = np.abs(gexp - oexp).max(1) / np.abs(oexp).max(1)
print("exp_spd device vs oracle: max rel", e2.max(), "n>1e-6:", int((e2 > 1e-6).sum()))
for i in np.argsort(e2)[-3:]:
    print(" worst", i, e2[i], "\n in", olog[i].reshape(3, 3), "\n gpu", gexp[i].reshape(3, 3), "\n orc", oexp[i].reshape(3, 3))
# 2. the map
for pose in (np.eye(4, dtype=np.float32), oracle.se3_exp(np.array([0.02, -0.01, 0.03, 1.0, 0.5, -0.2], np.float32))):
    gm, om = spx.VoxelHashMap(q, 0.5), oracle.VoxelHashMap(0.5)
    gm.add_point_cloud(cloud, pose)
    om.add_point_cloud(pts, pose, c16)
    res, keys = gm.downsampling(None, (0, 0, 0), 1e4, return_keys=True)
    want = om.downsampling((0, 0, 0), 1e4)
    o1, o2 = np.argsort(keys), np.argsort(want["keys"])
    assert np.array_equal(keys[o1], want["keys"][o2])
    gc, wc = res.covs.download(res.size())[o1], want["covs"][o2]
    # counts per voxel from the oracle's keys
    inv = np.float32(1.0) / np.float32(0.5)
    allk = np.array([oracle.voxel_key(oracle.transform_points(pose, pts[i:i + 1])[0], float(inv)) for i in range(n)], np.uint64)
    uk, cnt = np.unique(allk, return_counts=True)
    cnt_of = dict(zip(uk.tolist(), cnt.tolist()))
    counts = np.array([cnt_of[int(k)] for k in keys[o1]])
    rel = np.abs(gc - wc).max(1) / np.abs(wc).max(1)
    for c in (1, 2, 3, 4):
        sel = counts == c if c < 4 else counts >= 4
        if sel.any():
            print(f"count {c}{'+' if c == 4 else ''}: voxels {int(sel.sum())}, rel err max {rel[sel].max():.3e}, "
                  f"p99 {np.percentile(rel[sel], 99):.3e}, exact {int((rel[sel] == 0).sum())}")
    for i in np.argsort(rel)[-3:]:
        print(" worst voxel count", counts[i], rel[i], "\n gpu", gc[i].reshape(4, 4).T[:3, :3], "\n orc", wc[i].reshape(4, 4).T[:3, :3])
