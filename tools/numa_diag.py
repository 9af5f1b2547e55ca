"""Host-side placement diagnostics for the end-to-end (host buffers) leg: NUMA nodes visible to this process, the
GPU's own node, and the H2D bandwidth of a pinned 49 MB buffer allocated while the thread is bound to each node's
CPUs.  usage: python tools/numa_diag.py"""
import glob
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sycl_points_b200 as spx  # noqa: E402


def parse_cpulist(s):
    out = set()
    for part in s.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out.update(range(int(a), int(b or a) + 1))
    return out


def main():
    allowed = os.sched_getaffinity(0)
    print("cpu_count", os.cpu_count(), "affinity", sorted(allowed))
    nodes = {}
    for d in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
        nodes[int(d.rsplit("node", 1)[1])] = parse_cpulist(open(d + "/cpulist").read())
    for k, v in nodes.items():
        print(f"node {k}: {len(v)} cpus, {len(v & allowed)} allowed: {sorted(v & allowed)[:8]}...")
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=index,pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True).stdout.strip().splitlines()
        for line in bus:
            idx, bid = [x.strip() for x in line.split(",")]
            path = "/sys/bus/pci/devices/" + bid[-12:].lower() + "/numa_node"
            print("gpu", idx, bid, "numa_node", open(path).read().strip() if os.path.exists(path) else "?")
    except Exception as e:  # noqa: BLE001
        print("nvidia-smi:", e)
    q = spx.DeviceQueue(0)
    nbytes = 49 << 20
    dev = spx.DeviceArray(q, (nbytes,), np.uint8)

    def measure(tag):
        pin = spx.PinnedArray((nbytes,), np.uint8)
        pin.array[...] = 1
        a, b = spx.Event(), spx.Event()
        ts = []
        for _ in range(30):
            a.record(q)
            spx._lib.check(spx.lib().spx_memcpy_h2d(q.handle, dev.ptr, pin.array.ctypes.data, nbytes))
            b.record(q)
            q.wait()
            ts.append(a.elapsed_ms(b))
        ts = np.array(ts[3:])
        print(f"{tag}: H2D {nbytes / np.median(ts) / 1e6:.1f} GB/s median, {nbytes / ts.max() / 1e6:.1f} worst, "
              f"{nbytes / ts.min() / 1e6:.1f} best")

    measure("default placement")
    for k, cpus in nodes.items():
        use = cpus & allowed
        if not use:
            continue
        os.sched_setaffinity(0, use)
        measure(f"allocated bound to node {k}")
    os.sched_setaffinity(0, allowed)
    # packed upload the way bench.py's e2e leg does it: time per pair over 40 back-to-back copies
    pin = spx.PinnedArray((nbytes,), np.uint8)
    pin.array[...] = 1
    q.wait()
    t0 = time.perf_counter()
    for _ in range(40):
        spx._lib.check(spx.lib().spx_memcpy_h2d(q.handle, dev.ptr, pin.array.ctypes.data, nbytes))
    q.wait()
    dt = (time.perf_counter() - t0) / 40
    print(f"back-to-back: {dt * 1e3:.3f} ms per 49 MiB = {nbytes / dt / 1e9:.1f} GB/s")


if __name__ == "__main__":
    main()
