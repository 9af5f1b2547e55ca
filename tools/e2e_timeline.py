"""Timeline of the streamed (host buffers) pipeline of bench.py's pair workload: per step, when the upload of the
pair landed, when its two feeder chains finished and when its align returned, all relative to one base event;
plus host wall-clock marks of the issuing thread.  usage: python tools/e2e_timeline.py [steps]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import sycl_points_b200 as spx  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    streamed = os.environ.get("RESIDENT") is None
    q = spx.DeviceQueue(0)
    pairs = bench.rotating_pairs(0)
    pipe = bench.PairPipeline(spx, q, [len(p[1]) for p in pairs], [len(p[0]) for p in pairs])
    pins = []
    for tgt_raw, src_raw, _ in pairs:
        ps, pt = spx.PinnedArray((len(src_raw), 3)), spx.PinnedArray((len(tgt_raw), 3))
        ps.array[...] = src_raw[:, :3]
        pt.array[...] = tgt_raw[:, :3]
        pins.append((ps, pt))
        pipe.add_resident(src_raw, tgt_raw)
    if os.environ.get("PREGROW"):
        gb = int(os.environ["PREGROW"])
        for qq in (q, pipe.q2, pipe.qc, pipe.q3):
            tmp = spx.DeviceArray(qq, (gb << 30,), np.uint8)
            del tmp
            qq.wait()
    for w in range(5):
        pipe.run(w % len(pairs))
    NR = len(pairs)

    def sync_all():
        for qq in (q, pipe.q2, pipe.qc, pipe.q3):
            qq.wait()

    def run(n_steps, log):
        base, done = spx.Event(), spx.Event()
        sync_all()
        base.record(q)
        for qq in (pipe.q2, pipe.qc, pipe.q3):
            qq.wait_event(base)
        t_base = time.perf_counter()
        S = bench.STREAM_SLOTS
        if streamed:
            pipe.stream_upload(0, pins[0][0].array, pins[0][1].array)
            pipe.stream_upload(1, pins[1 % NR][0].array, pins[1 % NR][1].array)
        futs = pipe.feeders_async(0, 0, base, streamed)
        for s in range(n_steps):
            h0 = time.perf_counter()
            nxt = None
            if streamed and s + 2 < n_steps:
                pipe.stream_upload((s + 2) % S, pins[(s + 2) % NR][0].array, pins[(s + 2) % NR][1].array)
            if s + 1 < n_steps:
                if streamed:
                    nxt = pipe.feeders_async((s + 1) % S, (s + 1) % 2, None, True)
                else:
                    nxt = pipe.feeders_async((s + 1) % NR, (s + 1) % 2, None, False)
            h1 = time.perf_counter()
            src, tree_s = futs[0].result()
            tgt, tree_t = futs[1].result()
            hf = time.perf_counter()
            ds, dt_ = pipe.done_pipe[s % 2]
            pipe.q3.wait_event(ds)
            pipe.q3.wait_event(dt_)
            pipe.reg3.align(src, tgt, tree_t)
            ha = time.perf_counter()
            tree_s.close()
            tree_t.close()
            done.record(pipe.q3)
            pipe.q3.wait()
            h2 = time.perf_counter()
            row = dict(step=s, host_issue_ms=(h1 - h0) * 1e3, host_align_ms=(h2 - h1) * 1e3, host_t=(h2 - t_base) * 1e3,
                       host_futs_ms=(hf - h1) * 1e3, host_alignfn_ms=(ha - hf) * 1e3,
                       src_feed=base.elapsed_ms(ds), tgt_feed=base.elapsed_ms(dt_), align=base.elapsed_ms(done))
            if streamed and s > 0:
                pass
            log.append(row)
            futs = nxt
        sync_all()

    run(6, [])
    if os.environ.get("NOGC"):
        import gc
        gc.collect()
        gc.freeze()
        gc.disable()
    log = []
    run(steps, log)
    prev = 0.0
    if steps > 40:  # long run: only the outliers
        al = np.array([r["align"] for r in log])
        per = np.diff(al)
        med = float(np.median(per))
        print("mode:", "streamed" if streamed else "resident", "steps", steps, "median period %.3f ms" % med,
              "mean %.3f ms" % per.mean(), "-> %.0f pairs/s" % (1e3 / per.mean()), "gc", "off" if os.environ.get("NOGC") else "on")
        for i, p_ in enumerate(per):
            if p_ > 2.0 * med + 0.5:
                r = log[i + 1]
                print(f"  outlier step {r['step']}: period {p_:.3f} ms; feeders done {r['src_feed'] - al[i]:.3f} / "
                      f"{r['tgt_feed'] - al[i]:.3f} after prev align; host: issue {r['host_issue_ms']:.3f}, wait futures "
                      f"{r['host_futs_ms']:.3f}, align call {r['host_alignfn_ms']:.3f}")
        return
    print("mode:", "streamed (H2D inside)" if streamed else "resident")
    print(" step | feeders done src / tgt (ms) | align done | period | host: issue next / wait+align")
    for r in log:
        print(f"{r['step']:5d} | {r['src_feed']:8.3f} {r['tgt_feed']:8.3f} | {r['align']:8.3f} | {r['align'] - prev:6.3f} | "
              f"{r['host_issue_ms']:6.3f} {r['host_align_ms']:6.3f}")
        prev = r["align"]
    per = np.diff([r["align"] for r in log])[4:]
    print(f"steady period {per.mean():.3f} ms (min {per.min():.3f}, max {per.max():.3f}) -> {1e3 / per.mean():.0f} pairs/s")


if __name__ == "__main__":
    main()
