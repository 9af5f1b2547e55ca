"""BASELINE config 5 building block: P pairs of ~60 k points aligned by ONE batched launch
(spx_registration_align_batch) vs one align after the other.  usage: python tools/bench_batch.py [P] [reps]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import synthetic  # noqa: E402
import sycl_points_b200 as spx  # noqa: E402


def make_pairs(q, P, voxel=0.25, scenes=4):
    rs = np.random.RandomState(5)
    vg = spx.VoxelGrid(q, voxel)
    sc = []
    for seed in range(scenes):
        tgt_raw, _, _ = synthetic.kitti_pair(100 + seed, sweeps=8)
        tgt = vg.downsampling(spx.PointCloudShared(q, tgt_raw))
        tree = spx.KDTree.build(q, tgt)
        spx.covariance.estimate(tree.knn_search(tgt, 10), tgt)
        sc.append((tgt, tree, tgt_raw))
    pairs = []
    for j in range(P):
        tgt, tree, tgt_raw = sc[j % scenes]
        T = synthetic.random_pose(rs, 0.6, 1.0)
        keep = tgt_raw[rs.rand(len(tgt_raw)) < 0.9].astype(np.float64)
        src_raw = (keep @ np.linalg.inv(T).T).astype(np.float32)
        src_raw[:, 3] = 1.0
        src = vg.downsampling(spx.PointCloudShared(q, src_raw))
        ts = spx.KDTree.build(q, src)
        spx.covariance.estimate(ts.knn_search(src, 10), src)
        ts.close()
        pairs.append((src, tgt, tree, None))
    return pairs


def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    q = spx.DeviceQueue(0)
    pairs = make_pairs(q, P)
    params = spx.RegistrationParams()
    params.robust.type = spx.RobustLossType.HUBER
    reg = spx.Registration(q, params)
    ns = sum(p[0].size() for p in pairs)
    nt = sum(p[1].size() for p in pairs)
    for mode in ("batch", "single"):
        ms_all = []
        for r in range(reps + 1):
            a, b = spx.Event(), spx.Event()
            q.wait()
            a.record(q)
            if mode == "batch":
                res = reg.align_batch(pairs)
                kern_ms = reg.last_timing()["loop_ms"]
            else:
                res = [reg.align(*p) for p in pairs]
                kern_ms = float("nan")
            b.record(q)
            ms = a.elapsed_ms(b)
            if r:
                ms_all.append((ms, kern_ms))
        ms, kern = np.median([m[0] for m in ms_all]), np.median([m[1] for m in ms_all])
        its = [r.iterations + 1 for r in res]
        point_iters = sum(p[0].size() * i for p, i in zip(pairs, its))
        alg = 192 * point_iters + 16 * nt  # 192 B per source point and iteration + the index payload once
        print(f"{mode:6s} P={P} ns/pair={ns // P} total {ms:.3f} ms ({kern:.3f} kernel) -> {P / ms * 1e3:.0f} pairs/s, "
              f"iters mean {np.mean(its):.1f} max {max(its)}, {ms * 1e3 / sum(its):.1f} us per pair-iteration, "
              f"algorithmic {alg / 1e6:.0f} MB -> {alg / (kern if mode == 'batch' else ms) / 1e6:.0f} GB/s")


if __name__ == "__main__":
    main()
