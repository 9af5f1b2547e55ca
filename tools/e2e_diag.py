"""Why is the streamed end-to-end step slower than the device-resident one?  Times, on the bench
workload: the H2D copy alone, the compute loop alone (pre-uploaded buffers), and both overlapped."""
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
pin_src, pin_tgt = spx.PinnedArray(src_raw.shape), spx.PinnedArray(tgt_raw.shape)
pin_src.array[...] = src_raw
pin_tgt.array[...] = tgt_raw
K = 20


def sync_all():
    q.wait(); pipe.q2.wait(); pipe.qc.wait()


for slot in (0, 1):
    pipe.stream_upload(slot, pin_src.array, pin_tgt.array)
sync_all()
for _ in range(3):
    pipe.run_streamed(0)
a, b = spx.Event(), spx.Event()
sync_all()
a.record(pipe.qc)
for _ in range(K):
    pipe.stream_upload(0, pin_src.array, pin_tgt.array)
b.record(pipe.qc)
sync_all()
ms = a.elapsed_ms(b) / K
print(f"H2D alone: {ms:.3f} ms per pair = {(src_raw.nbytes + tgt_raw.nbytes) / ms / 1e6:.1f} GB/s")
sync_all()
t0 = time.perf_counter()
a.record(q)
for s in range(K):
    pipe.run_streamed(s % 2)
b.record(q)
sync_all()
print(f"compute alone (buffers resident): {a.elapsed_ms(b) / K:.3f} ms per pair (wall {(time.perf_counter() - t0) / K * 1e3:.3f})")
sync_all()
a.record(q)
pipe.qc.wait_event(a)
pipe.stream_upload(0, pin_src.array, pin_tgt.array)
t0 = time.perf_counter()
for s in range(K):
    if s + 1 < K:
        pipe.stream_upload((s + 1) % 2, pin_src.array, pin_tgt.array)
    pipe.run_streamed(s % 2)
b.record(q)
sync_all()
print(f"streamed (copy s+1 || compute s): {a.elapsed_ms(b) / K:.3f} ms per pair (wall {(time.perf_counter() - t0) / K * 1e3:.3f})")
# host cost of issuing the uploads
t0 = time.perf_counter()
for s in range(K):
    pipe.stream_upload(s % 2, pin_src.array, pin_tgt.array)
t1 = time.perf_counter()
sync_all()
print(f"host time to ISSUE one pair's uploads: {(t1 - t0) / K * 1e3:.3f} ms")
