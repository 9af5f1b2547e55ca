"""Host-side logic of the multi-GPU path on CPU: world_size-2 `gloo` process group, the per-shard
linearisation injected (the oracle stands in for the CUDA kernels, as the reference's own tests
inject KNNs: T/test_registration_pipeline.cpp:16-61).  Covers shard partitioning, the
reduce-then-identical-update protocol (every rank ends with the same pose, equal to the unsharded
align) and bench.py's reference arm under torchrun (rank 0 works, the others exit 0)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_bounds_cover_and_are_contiguous():
    from sycl_points_b200.multi_gpu import shard_bounds, shard_of
    for n in (0, 1, 7, 8, 9, 1000, 114045):
        for w in (1, 2, 3, 4, 8):
            b = shard_bounds(n, w)
            assert len(b) == w and b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert all(0 <= hi - lo <= -(-n // w) for lo, hi in b)
            assert shard_of(n, w - 1, w) == b[-1]
    with pytest.raises(ValueError):
        shard_bounds(10, 0)


def test_sums_row_layout_roundtrip():
    from sycl_points_b200.multi_gpu import S_B, S_ERR, S_INL, SUMS_LEN, sums_to_Hb
    row = np.zeros(SUMS_LEN)
    row[:21] = np.arange(1, 22)
    row[S_B:S_B + 6] = [-1, -2, -3, -4, -5, -6]
    row[S_ERR], row[S_INL] = 2.5, 1234
    H, b, e, inl = sums_to_Hb(row)
    assert np.array_equal(H, H.T) and H[0, 0] == 1 and H[0, 5] == 6 and H[1, 1] == 7 and H[5, 5] == 21
    assert np.array_equal(b, [-1, -2, -3, -4, -5, -6]) and e == 2.5 and inl == 1234


WORKER = r'''
import os, sys, json
import numpy as np
sys.path.insert(0, sys.argv[1])
import torch
import torch.distributed as dist
import oracle
from sycl_points_b200.multi_gpu import ShardedAlignLoop, shard_of, sums_to_Hb, S_B, S_ERR, S_INL, SUMS_LEN

dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=int(sys.argv[3]), world_size=int(sys.argv[4]))
rank, world = dist.get_rank(), dist.get_world_size()
d = np.load(os.path.join(sys.argv[1], "tests", "golden", "bundled_pair.npz"))
src, tgt = d["source_ds"], d["target_ds"]
tree_s, tree_t = oracle.KDTree(src), oracle.KDTree(tgt)
cs = oracle.covariance(src, tree_s.knn(src, 10)[0])
ct = oracle.covariance(tgt, tree_t.knn(tgt, 10)[0])
REG, LOSS, SCALE, MAXC, LAM, CRIT = 3, 1, 10.0, 2.0, 1.0, 1e-3
lo, hi = shard_of(len(src), rank, world)


class OracleShardLoop(ShardedAlignLoop):
    def begin(self):
        self.T = np.eye(4, dtype=np.float32)
        self.iterations, self.converged = 0, False

    def linearize_shard(self):
        idx, dist_ = tree_t.knn(src[lo:hi], 1, T=self.T)
        H, b, e, inl = oracle.linearize(REG, LOSS, src[lo:hi], cs[lo:hi], tgt, ct, None, idx, dist_, self.T,
                                        MAXC * MAXC, SCALE, mode=1)
        row = np.zeros(SUMS_LEN)
        t = 0
        for a in range(6):
            for c in range(a, 6):
                row[t] = H[a, c]
                t += 1
        row[S_B:S_B + 6] = b
        row[S_ERR], row[S_INL] = e, inl
        return torch.from_numpy(row)

    def update(self, row):
        H, b, e, inl = sums_to_Hb(row.numpy())
        ok, delta = oracle.solve6(H, b, LAM)
        self.T = (self.T.astype(np.float32) @ oracle.se3_exp(delta)).astype(np.float32)
        self.converged = bool(ok and np.linalg.norm(delta[:3]) < CRIT and np.linalg.norm(delta[3:]) < CRIT)
        self.inlier = inl
        done = self.converged
        self.iterations += 0 if done else 1
        return done

    def finish(self):
        return dict(T=self.T, iterations=self.iterations, converged=self.converged, inlier=self.inlier)


out = OracleShardLoop(20, lambda t: dist.all_reduce(t)).run()
# every rank must hold the same pose bit for bit (identical reduced sums -> identical update)
gathered = [None] * world
dist.all_gather_object(gathered, out["T"].tobytes())
assert all(g == gathered[0] for g in gathered), "ranks diverged"
if rank == 0:
    P = oracle.default_params(reg_type=REG, loss=LOSS, robust_default_scale=SCALE)
    ref = oracle.align(P, src, cs, tgt, ct, None, tree_t)
    dT = np.linalg.inv(ref["T"].astype(np.float64)) @ out["T"].astype(np.float64)
    ang = float(np.linalg.norm(0.5 * np.array([dT[2, 1] - dT[1, 2], dT[0, 2] - dT[2, 0], dT[1, 0] - dT[0, 1]])))
    print(json.dumps(dict(dt=float(np.linalg.norm(dT[:3, 3])), ang=ang, it=out["iterations"], ref_it=ref["iterations"],
                          conv=out["converged"], ref_conv=ref["converged"], inl=out["inlier"], ref_inl=ref["inlier"],
                          shard=[lo, hi])))
dist.destroy_process_group()
'''


@pytest.mark.timeout(300)
def test_sharded_loop_world2_gloo_matches_unsharded(tmp_path):
    """2 ranks x half the source each == the unsharded oracle align (pose <= 1e-5 m / 1e-5 rad)."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = free_port()
    env = dict(os.environ, OMP_NUM_THREADS="2")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r), "2"], stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True, env=env) for r in range(2)]
    outs = [p.communicate(timeout=280) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-2000:]
    import json
    res = json.loads(outs[0][0].strip().splitlines()[-1])
    assert res["dt"] < 1e-5 and res["ang"] < 1e-5, res
    assert res["it"] == res["ref_it"] and res["conv"] == res["ref_conv"] and res["inl"] == res["ref_inl"], res


@pytest.mark.timeout(600)
def test_bench_reference_arm_under_torchrun_rank0_only():
    """bench.py --impl reference with WORLD_SIZE=2: rank 0 prints the JSON line, rank 1 exits 0 silently."""
    env = dict(os.environ, WORLD_SIZE="2", LOCAL_RANK="1", RANK="1", SPX_BENCH_TINY="1")
    r1 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                         "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env, timeout=120)
    assert r1.returncode == 0 and r1.stdout.strip() == "", (r1.stdout, r1.stderr[-500:])


def test_block_cyclic_shards_partition_the_cloud():
    from sycl_points_b200.multi_gpu import shard_indices
    for n in (0, 1, 1023, 1024, 5000, 114045):
        for w in (1, 2, 3, 8):
            parts = [shard_indices(n, r, w) for r in range(w)]
            allidx = np.sort(np.concatenate(parts)) if parts else np.zeros(0)
            assert np.array_equal(allidx, np.arange(n))
            if n >= 8 * 1024 * w:
                sizes = [len(p) for p in parts]
                assert max(sizes) - min(sizes) <= 1024


WORKER_KNN = r'''
import os, sys, json
import numpy as np
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
import oracle
from sycl_points_b200.multi_gpu import ShardedKNN, ShardedCovariance, shard_of

dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=int(sys.argv[3]), world_size=int(sys.argv[4]))
rank, world = dist.get_rank(), dist.get_world_size()
Q = oracle.Rng(1234).box_points(1001, (-50, -50, -3), (50, 50, 10))   # config-3 distribution, odd count: ragged shards
T = oracle.Rng(4321).box_points(3000, (-50, -50, -3), (50, 50, 10))
knn = ShardedKNN(rank, world, search=lambda rows, k: oracle.knn_bruteforce(rows, T, k))
idx, dst, (lo, hi) = knn.knn_search(Q, 20)
assert (lo, hi) == shard_of(len(Q), rank, world) and idx.shape == (hi - lo, 20)
gi, gd, _ = knn.knn_search(Q, 20, gather=True)
ri, rd = oracle.knn_bruteforce(Q, T, 20)
ok_knn = bool(np.array_equal(gi, ri) and np.array_equal(gd, rd) and np.array_equal(idx, ri[lo:hi]))
# covariance: every rank holds the full cloud, computes its rows, all-gathers
tree = oracle.KDTree(T)
def compute(points, lo, hi, k):
    nn, _ = tree.knn(points[lo:hi], k)
    # rows of the shard against the FULL cloud: estimate() gathers neighbours from `points`
    full_idx = np.full((len(points), k), -1, np.int32)
    full_idx[lo:hi] = nn
    return oracle.covariance(points, full_idx)[lo:hi]
cov, _ = ShardedCovariance(rank, world, compute=compute).estimate(T, 10)
ref = oracle.covariance(T, tree.knn(T, 10)[0])
ok_cov = bool(np.array_equal(cov, ref))
# an empty query set and more ranks than rows
e_i, e_d, _ = knn.knn_search(Q[:0], 5, gather=True)
o_i, o_d, _ = knn.knn_search(Q[:1], 5, gather=True)
ok_edge = bool(e_i.shape == (0, 5) and np.array_equal(o_i, ri[:1, :5]))
if rank == 0:
    print(json.dumps(dict(knn=ok_knn, cov=ok_cov, edge=ok_edge)))
dist.destroy_process_group()
'''


@pytest.mark.timeout(300)
def test_sharded_knn_and_covariance_world2_gloo(tmp_path):
    """SURVEY §8(e) rows 2-3 on CPU: queries / points split over 2 gloo ranks, per-shard computation injected
    (oracle); the gathered result equals the unsharded one bit for bit, ragged and empty shards included."""
    script = tmp_path / "worker_knn.py"
    script.write_text(WORKER_KNN)
    port = free_port()
    env = dict(os.environ, OMP_NUM_THREADS="2")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r), "2"], stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True, env=env) for r in range(2)]
    outs = [p.communicate(timeout=280) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-2000:]
    import json
    res = json.loads(outs[0][0].strip().splitlines()[-1])
    assert res == dict(knn=True, cov=True, edge=True), res
