"""BASELINE config 4: point-to-plane + GICP on a dense (~2 M points after a 0.05 m voxel grid) scan
pair with the SOURCE SHARDED across the ranks and the target + index replicated.

    python tools/bench_sharded.py [--iters 20] [--small]                       (1 GPU)
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_sharded.py                                                 (N GPUs, one rank each)

Per factor it times `--iters` forced Gauss-Newton iterations (convergence criteria 0) three ways:
the plain single-GPU align on the whole source (rank 0), the sharded align with the in-kernel
NVLink exchange ("p2p"), and the sharded align with one NCCL all-reduce per iteration ("nccl").
Device time by CUDA events on the launching stream, max over ranks.  Parity: every rank's sharded
pose must equal the single-GPU pose within 1e-5 m / 1e-5 rad.  One JSON line per case (rank 0).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import synthetic  # noqa: E402


def pose_delta(Ta, Tb):
    d = np.linalg.inv(Ta.astype(np.float64)) @ Tb.astype(np.float64)
    w = 0.5 * np.array([d[2, 1] - d[1, 2], d[0, 2] - d[2, 0], d[1, 0] - d[0, 1]])
    return float(np.linalg.norm(d[:3, 3])), float(np.arcsin(min(1.0, np.linalg.norm(w))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--small", action="store_true", help="config-2 sized pair (quick check)")
    ap.add_argument("--single-only", action="store_true", help="only the plain single-GPU align (ncu captures)")
    ap.add_argument("--factors", default="POINT_TO_PLANE,GICP")
    ap.add_argument("--contiguous", action="store_true", help="contiguous ceil(N/G) shards instead of block-cyclic")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import sycl_points_b200 as spx
    from sycl_points_b200.multi_gpu import Communicator, ShardedRegistration, shard_indices, shard_of

    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        q = spx.DeviceQueue(local, cuda_stream=stream.cuda_stream)
        t0 = time.time()
        if args.small:
            tgt_raw, src_raw, T_gt = synthetic.kitti_pair(42)
            voxel, cap = 0.25, 10**9
        else:
            tgt_raw, src_raw, T_gt = synthetic.dense_pair(42)
            voxel, cap = 0.05, 2_000_000
        vg = spx.VoxelGrid(q, voxel)
        src_full = vg.downsampling(spx.PointCloudShared(q, src_raw)).points_host()[:cap]
        tgt_full = vg.downsampling(spx.PointCloudShared(q, tgt_raw)).points_host()[:cap]
        del src_raw, tgt_raw
        src, tgt = spx.PointCloudShared(q, src_full), spx.PointCloudShared(q, tgt_full)
        ts, tt = spx.KDTree.build(q, src), spx.KDTree.build(q, tgt)
        nn_s, nn_t = ts.knn_search(src, 10), tt.knn_search(tgt, 10)
        spx.covariance.estimate(nn_s, src)
        spx.covariance.estimate(nn_t, tgt)
        spx.covariance.estimate_normals(nn_t, tgt)
        ts.close()
        ns, nt = src.size(), tgt.size()
        lo, hi = shard_of(ns, rank, world)
        sel = np.arange(lo, hi) if args.contiguous else shard_indices(ns, rank, world)
        cov_s = src.covs_host()
        shard = spx.PointCloudShared(q, src_full[sel], cov_s[sel])
        if rank == 0:
            print(f"# setup {time.time() - t0:.1f} s: N_s={ns} N_t={nt} world={world} shard0={len(sel)} pts ({'contiguous' if args.contiguous else 'block-cyclic'}) index={tt.info()}",
                  file=sys.stderr)
        comm = Communicator(q, rank, world) if world > 1 else Communicator(q, 0, 1)
        peak = 6546.6
        try:
            peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            pass

        def maxed(ms):
            if world == 1:
                return ms
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        def barrier():
            q.wait()
            if world > 1:
                dist.barrier()
                torch.cuda.synchronize()

        for regname, per_pt in (("POINT_TO_PLANE", 80), ("GICP", 192)):
            if regname not in args.factors.split(","):
                continue
            params = spx.RegistrationParams(reg_type=spx.RegType[regname], max_iterations=args.iters)
            params.robust.type = spx.RobustLossType.HUBER
            params.criteria.translation = params.criteria.rotation = 0.0
            alg_bytes = per_pt * ns + 16 * nt  # per iteration, whole job
            results = {}
            # ---- single GPU, whole source (every rank runs it: same clocks everywhere; rank 0 reports)
            reg1 = spx.Registration(q, params)
            ref = reg1.align(src, tgt, tt)
            ms = []
            for _ in range(args.reps):
                barrier()
                reg1.align(src, tgt, tt)
                ms.append(reg1.last_timing()["loop_ms"])
            results["single"] = (float(np.median(ms)), ref)
            # ---- sharded
            modes = [] if args.single_only else ["p2p"] + (["nccl"] if world > 1 else [])
            for mode in modes:
                sreg = ShardedRegistration(q, params, comm=comm, mode=mode)
                out = sreg.align(shard, tgt, tt)
                ms = []
                for _ in range(args.reps):
                    barrier()
                    a, b = spx.Event(), spx.Event()
                    a.record(q)
                    out = sreg.align(shard, tgt, tt)
                    b.record(q)
                    ms.append(maxed(a.elapsed_ms(b) if mode == "nccl" else sreg.last_timing()["loop_ms"]))
                results[mode] = (float(np.median(ms)), out)
            if rank == 0:
                for mode, (t_ms, out) in results.items():
                    dt, da = pose_delta(ref.T, out.T)
                    n_g = 1 if mode == "single" else world
                    per_iter = t_ms / args.iters
                    print(json.dumps({
                        "workload": "config 4 dense pair" if not args.small else "config 2 pair (small)",
                        "factor": regname, "mode": mode, "n_gpus": n_g, "N_s": ns, "N_t": nt, "iterations": args.iters,
                        "ms_per_iter": per_iter, "align_ms": t_ms,
                        "algorithmic_bytes_per_iter": alg_bytes, "achieved_GBps": alg_bytes / (per_iter * 1e-3) / 1e9,
                        "frac_of_hbm_peak_per_gpu": alg_bytes / (per_iter * 1e-3) / 1e9 / (peak * n_g),
                        "pose_vs_single_m": dt, "pose_vs_single_rad": da, "inlier": int(out.inlier),
                        "pose_err_vs_gt_m": pose_delta(T_gt, out.T)[0]}), flush=True)
                    if mode != "single":
                        assert dt < 1e-5 and da < 1e-5, (mode, dt, da)
        barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
