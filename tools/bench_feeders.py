"""Throughput of the per-cloud feeder chain (voxel grid -> index build -> KNN k=10 -> covariance) when W host
threads drive W queues concurrently.  usage: python tools/bench_feeders.py [clouds] [sweeps]"""
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import synthetic  # noqa: E402
import sycl_points_b200 as spx  # noqa: E402


def main():
    n_clouds = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    q0 = spx.DeviceQueue(0)
    raws = []
    for seed in range(4):
        tgt_raw, src_raw, _ = synthetic.kitti_pair(100 + seed, sweeps=sweeps)
        raws += [tgt_raw, src_raw]
    clouds = [spx.PointCloudShared(q0, raws[j % len(raws)]) for j in range(n_clouds)]
    q0.wait()
    print("raw points per cloud", len(raws[0]))
    for W in (1, 2, 4, 8, 16):
        qs = [spx.DeviceQueue(0) for _ in range(W)]
        for qq in qs:
            qq.set_blocking_sync(W > 8)
        out = [None] * n_clouds

        def work(w):
            q = qs[w]
            vg = spx.VoxelGrid(q, 0.25)
            nn = spx.KNNResult()
            for j in range(w, n_clouds, W):
                raw = clouds[j]
                c = spx.PointCloudShared(q)
                c.adopt_points(raw.points, raw.size())
                ds = vg.downsampling(c)
                tree = spx.KDTree.build(q, ds)
                tree.knn_search_async(ds, 10, nn)
                spx.covariance.estimate(nn, ds)
                q.wait()
                out[j] = ds.size()
                tree.close()

        for rep in range(3):
            t0 = time.perf_counter()
            th = [threading.Thread(target=work, args=(w,)) for w in range(W)]
            for t in th:
                t.start()
            for t in th:
                t.join()
            dt = time.perf_counter() - t0
        print(f"W={W:2d}: {n_clouds} clouds in {dt * 1e3:.2f} ms -> {dt / n_clouds * 1e6:.0f} us per cloud, out {out[0]}")


if __name__ == "__main__":
    main()
