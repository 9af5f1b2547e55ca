"""BASELINE config 3: brute-force KNN k=20, nq x nt uniform points, and the index
search on the same data; a sample of the queries is checked against the oracle-free property that
index == brute force.  usage: python tools/bench_knn.py [nq] [nt] [k]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sycl_points_b200 as spx  # noqa: E402


def cloud(n, seed):
    rs = np.random.RandomState(seed)
    p = np.empty((n, 4), np.float32)
    p[:, 0] = rs.uniform(-50, 50, n)
    p[:, 1] = rs.uniform(-50, 50, n)
    p[:, 2] = rs.uniform(-3, 10, n)
    p[:, 3] = 1.0
    return p


def main():
    nq = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    nt = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    k = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    q = spx.DeviceQueue(0)
    Q, T = spx.PointCloudShared(q, cloud(nq, 1234)), spx.PointCloudShared(q, cloud(nt, 4321))
    a, b = spx.Event(), spx.Event()
    res = spx.knn_search_bruteforce(q, Q, T, k)  # warm-up
    q.wait()
    a.record(q)
    res = spx.knn_search_bruteforce(q, Q, T, k)
    b.record(q)
    ms = a.elapsed_ms(b)
    pairs = nq * nt
    print(f"bruteforce k={k} {nq}x{nt}: {ms:.2f} ms  {nq / ms / 1e3:.3f} Mqueries/s  {pairs / ms / 1e6:.1f} Gpair/s  "
          f"frac of fp32-issue floor(7 instr/pair @ 37.2T lane-instr/s): {pairs * 7 / 37.2e12 / (ms * 1e-3):.3f}")
    tree = spx.KDTree.build(q, T)
    r2 = spx.KNNResult()
    tree.knn_search_async(Q, k, r2)
    q.wait()
    a.record(q)
    tree.knn_search_async(Q, k, r2)
    b.record(q)
    ms2 = a.elapsed_ms(b)
    print(f"index      k={k} {nq}x{nt}: {ms2:.2f} ms  {nq / ms2 / 1e3:.3f} Mqueries/s  info={tree.info()}")
    same_i = np.array_equal(res.indices_host(), r2.indices_host())
    same_d = np.array_equal(res.distances_host(), r2.distances_host())
    print("index == bruteforce (bit-exact idx, dist):", same_i, same_d)


if __name__ == "__main__":
    main()
