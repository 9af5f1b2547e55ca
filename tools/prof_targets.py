"""One instance of every hot kernel inside a cudaProfilerStart/Stop range (for `ncu --profile-from-start off`):
the config-2 feeder chain of one cloud (voxel grid, index build, KNN k=10, covariance), a single-pair GICP align,
a batched align of P pairs of ~64 k points, and a brute-force KNN tile scan (64 k x 1 M, k = 20).
usage: python tools/prof_targets.py [P]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import synthetic  # noqa: E402
import sycl_points_b200 as spx  # noqa: E402
from tools.bench_batch import make_pairs  # noqa: E402


def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    q = spx.DeviceQueue(0)
    tgt_raw, src_raw, _ = synthetic.kitti_pair(42)
    raw_s, raw_t = spx.PointCloudShared(q, src_raw), spx.PointCloudShared(q, tgt_raw)
    vg = spx.VoxelGrid(q, 0.25)
    params = spx.RegistrationParams()
    params.robust.type = spx.RobustLossType.HUBER
    reg = spx.Registration(q, params)
    pairs = make_pairs(q, P)
    Qh, Th = synthetic.knn_config3(65536, 1_000_000)
    Q, T = spx.PointCloudShared(q, Qh), spx.PointCloudShared(q, Th)
    nn = spx.KNNResult()

    def everything():
        src = vg.downsampling(raw_s)
        tgt = vg.downsampling(raw_t)
        ts, tt = spx.KDTree.build(q, src), spx.KDTree.build(q, tgt)
        ts.knn_search_async(src, 10, nn)
        spx.covariance.estimate(nn, src)
        tt.knn_search_async(tgt, 10, nn)
        spx.covariance.estimate(nn, tgt)
        r = reg.align(src, tgt, tt)
        rb = reg.align_batch(pairs)
        spx.knn_search_bruteforce(q, Q, T, 20)
        q.wait()
        ts.close()
        tt.close()
        return r, rb

    for _ in range(2):
        everything()
    spx._lib.check(spx.lib().spx_profiler_range(1))
    r, rb = everything()
    spx._lib.check(spx.lib().spx_profiler_range(0))
    print("single align iterations", r.iterations + 1, "batch iterations", [x.iterations + 1 for x in rb][:8])


if __name__ == "__main__":
    main()
