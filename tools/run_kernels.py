"""Exercise the hot kernels in isolation on the config-2 synthetic pair (for ncu captures and
quick CUDA-event timings).  usage: python tools/run_kernels.py [align|knn|voxel|index|all] [reps]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import synthetic  # noqa: E402
import sycl_points_b200 as spx  # noqa: E402


def timed(q, fn, reps):
    fn()
    q.wait()
    a, b = spx.Event(), spx.Event()
    a.record(q)
    for _ in range(reps):
        fn()
    b.record(q)
    return a.elapsed_ms(b) / reps


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    q = spx.DeviceQueue(0)
    tgt_raw, src_raw, T_gt = synthetic.kitti_pair(42)
    raw_s, raw_t = spx.PointCloudShared(q, src_raw), spx.PointCloudShared(q, tgt_raw)
    vg = spx.VoxelGrid(q, 0.25)
    if what in ("voxel", "all"):
        print("voxel_downsample 2.0M -> ~120k: %.3f ms" % timed(q, lambda: vg.downsampling(raw_t), reps))
    src, tgt = vg.downsampling(raw_s), vg.downsampling(raw_t)
    print("points", src.size(), tgt.size())
    if what in ("index", "all"):
        def build():
            t = spx.KDTree.build(q, tgt)
            t.close()
        print("index build: %.3f ms" % timed(q, build, reps))
    ts, tt = spx.KDTree.build(q, src), spx.KDTree.build(q, tgt)
    print("index", tt.info())
    nn_s, nn_t = spx.KNNResult(), spx.KNNResult()
    if what in ("knn", "all"):
        print("knn k=10 self (%d q): %.3f ms" % (tgt.size(), timed(q, lambda: tt.knn_search_async(tgt, 10, nn_t), reps)))
        nn1 = spx.KNNResult()
        print("knn k=1 src->tgt: %.3f ms" % timed(q, lambda: tt.knn_search_async(src, 1, nn1), reps))
    ts.knn_search_async(src, 10, nn_s)
    tt.knn_search_async(tgt, 10, nn_t)
    if what in ("cov", "all"):
        print("covariance k=10: %.3f ms" % timed(q, lambda: spx.covariance.estimate(nn_t, tgt), reps))
    spx.covariance.estimate(nn_s, src)
    spx.covariance.estimate(nn_t, tgt)
    if what in ("align", "all"):
        for name, crit in (("default criteria", 1e-3), ("20 forced iterations", 0.0)):
            p = spx.RegistrationParams()
            p.robust.type = spx.RobustLossType.HUBER
            p.criteria.translation = p.criteria.rotation = crit
            reg = spx.Registration(q, p)
            t0 = time.perf_counter()
            ms = timed(q, lambda: reg.align(src, tgt, tt), reps)
            lt = reg.last_timing()
            print("align GICP (%s): %.3f ms/align, loop %.3f ms, %d iterations -> %.1f us/iter" %
                  (name, ms, lt["loop_ms"], lt["iterations"], 1e3 * lt["loop_ms"] / lt["iterations"]))
        for regt in ("POINT_TO_POINT", "POINT_TO_PLANE"):
            p = spx.RegistrationParams(reg_type=spx.RegType[regt])
            p.criteria.translation = p.criteria.rotation = 0.0
            if regt == "POINT_TO_PLANE":
                spx.covariance.estimate_normals(nn_t, tgt)
            reg = spx.Registration(q, p)
            timed(q, lambda: reg.align(src, tgt, tt), 2)
            lt = reg.last_timing()
            print("align %s: loop %.3f ms, %d iterations -> %.1f us/iter" %
                  (regt, lt["loop_ms"], lt["iterations"], 1e3 * lt["loop_ms"] / lt["iterations"]))


if __name__ == "__main__":
    main()
