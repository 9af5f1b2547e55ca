"""Timeline of one bench step (config-2 pair, two queues): CUDA-event offsets of every stage on both
queues relative to the step's start event, plus host wall-clock marks.  Shows where a step's time goes."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import synthetic  # noqa: E402
import sycl_points_b200 as spx  # noqa: E402

q = spx.DeviceQueue(0)
tgt_raw, src_raw, T_gt = synthetic.kitti_pair(42)
pipe = bench.PairPipeline(spx, q, len(src_raw), len(tgt_raw))
pipe.upload(src_raw, tgt_raw)
flush = spx.DeviceArray(q, (256 << 20,), np.uint8)
K = bench.K_COV


def chain(qq, vg, raw, nn, start, marks, host):
    qq.wait_event(start)
    ev = lambda: spx.Event().record(qq)  # noqa: E731
    host.append(time.perf_counter())
    cloud = vg.downsampling(raw); marks.append(("voxel", ev())); host.append(time.perf_counter())
    tree = spx.KDTree.build(qq, cloud); marks.append(("build", ev())); host.append(time.perf_counter())
    tree.knn_search_async(cloud, K, nn); marks.append(("knn", ev()))
    spx.covariance.estimate(nn, cloud); marks.append(("cov", ev())); host.append(time.perf_counter())
    return cloud, tree


for rep in range(6):
    spx._lib.check(spx.lib().spx_memset(q.handle, flush.ptr, 0, flush.nbytes))
    q.wait(); pipe.q2.wait()
    start = spx.Event().record(q)
    t0 = time.perf_counter()
    ms, mt, hs, ht = [], [], [], []
    fut = pipe.pool.submit(chain, pipe.q2, pipe.vg2, pipe.raw_src, pipe.nn_s, start, ms, hs)
    tgt, tt = chain(q, pipe.vg, pipe.raw_tgt, pipe.nn_t, start, mt, ht)
    src, ts = fut.result()
    t_join = time.perf_counter()
    q.wait_event(ms[-1][1])  # device-side join, as bench.py does
    a0 = spx.Event().record(q)
    res = pipe.reg.align(src, tgt, tt)
    a1 = spx.Event().record(q)
    t_end = time.perf_counter()
    q.wait()
    if rep < 3:
        continue
    off = lambda e: start.elapsed_ms(e)  # noqa: E731
    print(f"--- step {rep}: total {off(a1):.3f} ms (wall {1e3 * (t_end - t0):.3f})")
    print("  target queue: " + "  ".join(f"{n} {off(e):.3f}" for n, e in mt) + f"   host marks {[round(1e3 * (t - t0), 3) for t in ht]}")
    print("  source queue: " + "  ".join(f"{n} {off(e):.3f}" for n, e in ms) + f"   host marks {[round(1e3 * (t - t0), 3) for t in hs]}")
    print(f"  join at host {1e3 * (t_join - t0):.3f}; align gpu {off(a0):.3f} -> {off(a1):.3f}; kernel loop {pipe.reg.last_timing()['loop_ms']:.3f}")
    ts.close()
