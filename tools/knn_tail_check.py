import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np, synthetic, sycl_points_b200 as spx
q = spx.DeviceQueue(0)
tgt_raw, src_raw, T_gt = synthetic.kitti_pair(42)
vg = spx.VoxelGrid(q, 0.25)
src, tgt = vg.downsampling(spx.PointCloudShared(q, src_raw)), vg.downsampling(spx.PointCloudShared(q, tgt_raw))
for name, c in (("target", tgt), ("source", src)):
    t = spx.KDTree.build(q, c)
    nn = spx.KNNResult()
    t.knn_search_async(c, 10, nn); q.wait()
    a, b = spx.Event(), spx.Event()
    a.record(q)
    for _ in range(10): t.knn_search_async(c, 10, nn)
    b.record(q)
    d = nn.distances_host()
    kth = np.sqrt(d[:, 9])
    print(name, c.size(), "knn k=10: %.3f ms" % (a.elapsed_ms(b) / 10), t.info(), "10th-NN dist p50 %.2f p99 %.2f p99.9 %.2f max %.2f" % (np.percentile(kth, 50), np.percentile(kth, 99), np.percentile(kth, 99.9), kth.max()), "n(>2m)=", int((kth > 2).sum()), "n(>5m)=", int((kth > 5).sum()))
