"""Brute-force KNN tile scan: time every variant on config 3 (1 M x 1 M, k = 20); with `prof` as argument run ONE
320 k x 1 M launch of the variant in SPX_BF_VARIANT inside a profiler range (for ncu)."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import synthetic  # noqa: E402


def run_one(nq, nt, reps):
    import sycl_points_b200 as spx
    q = spx.DeviceQueue(0)
    Qh, Th = synthetic.knn_config3(nq, nt)
    Q, T = spx.PointCloudShared(q, Qh), spx.PointCloudShared(q, Th)
    spx.knn_search_bruteforce(q, Q, T, 20)
    q.wait()
    a, b = spx.Event(), spx.Event()
    best = 1e9
    for _ in range(reps):
        a.record(q)
        r = spx.knn_search_bruteforce(q, Q, T, 20)
        b.record(q)
        q.wait()
        best = min(best, a.elapsed_ms(b))
    return best, r


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "prof":
        import sycl_points_b200 as spx
        q = spx.DeviceQueue(0)
        Qh, Th = synthetic.knn_config3(327680, 1_000_000)
        Q, T = spx.PointCloudShared(q, Qh), spx.PointCloudShared(q, Th)
        spx.knn_search_bruteforce(q, Q, T, 20)
        q.wait()
        spx._lib.check(spx.lib().spx_profiler_range(1))
        spx.knn_search_bruteforce(q, Q, T, 20)
        q.wait()
        spx._lib.check(spx.lib().spx_profiler_range(0))
    elif len(sys.argv) > 1 and sys.argv[1] == "one":
        ms, r = run_one(1_000_000, 1_000_000, 2)
        print(f"{os.environ.get('SPX_BF_VARIANT', 'default'):8s} {ms:9.2f} ms  {1e3 / ms:6.2f} Mq/s   checksum {int(r.indices_host()[::997].sum())}")
    else:
        for v, b in (("p2", 1), ("p2", 2), ("p4", 1), ("p4", 2)):
            env = dict(os.environ, SPX_BF_VARIANT=v, SPX_BF_BATCH=str(b))
            print(f"batch {b} ", end="", flush=True)
            subprocess.run([sys.executable, __file__, "one"], env=env)
