"""Generate the committed fixtures under tests/golden/ from the reference's bundled scan pair.

Run HERE (the container that mounts /root/reference); the GPU box has no /root/reference, so
tests only ever read the .npz files this script writes.  Data licence: MIT (cpp/data/LICENSE,
Kenji Koide) — derived, down-sampled point sets only; no reference source code is copied.

    python tests/golden/make_fixtures.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

REF = "/root/reference/cpp/data"
OUT = os.path.dirname(os.path.abspath(__file__))


def read_ply_xyz(path):
    """Binary little-endian PLY with float x,y,z,scalar_intensity vertex records -> (n,4) xyz1."""
    with open(path, "rb") as f:
        n = None
        while True:
            line = f.readline().decode("ascii", "replace").strip()
            if line.startswith("element vertex"):
                n = int(line.split()[-1])
            if line == "end_header":
                break
        raw = np.frombuffer(f.read(n * 16), dtype="<f4").reshape(n, 4)
    pts = raw.copy()
    pts[:, 3] = 1.0
    return pts


def main():
    src = read_ply_xyz(os.path.join(REF, "source.ply"))
    tgt = read_ply_xyz(os.path.join(REF, "target.ply"))
    T_gt = np.loadtxt(os.path.join(REF, "T_target_source.txt")).astype(np.float32)
    print("raw", src.shape, tgt.shape)
    # example_registration.cpp:52-73: box filter 0.5..50 m then 0.25 m voxel grid
    src_b = oracle.box_filter(src, 0.5, 50.0)
    tgt_b = oracle.box_filter(tgt, 0.5, 50.0)
    src_ds = oracle.voxel_downsample(src_b, 0.25)
    tgt_ds = oracle.voxel_downsample(tgt_b, 0.25)
    print("box", len(src_b), len(tgt_b), "voxel", len(src_ds), len(tgt_ds))
    np.savez_compressed(os.path.join(OUT, "bundled_pair.npz"),
                        source_raw_head=src[:20000], source_ds=src_ds, target_ds=tgt_ds, T_target_source=T_gt,
                        counts=np.array([len(src), len(tgt), len(src_b), len(tgt_b), len(src_ds), len(tgt_ds)]))

    # goldens from the oracle on that pair (regression pins; generated here, checked everywhere)
    k = 10
    tree_s, tree_t = oracle.KDTree(src_ds), oracle.KDTree(tgt_ds)
    idx_s, _ = tree_s.knn(src_ds, k)
    idx_t, _ = tree_t.knn(tgt_ds, k)
    cov_s, cov_t = oracle.covariance(src_ds, idx_s), oracle.covariance(tgt_ds, idx_t)
    nrm_t = oracle.normals(tgt_ds, idx_t)
    nn_idx, nn_dist = tree_t.knn(src_ds, 1)
    gold = dict(idx_s_head=idx_s[:256], idx_t_head=idx_t[:256], cov_s_head=cov_s[:256], nrm_t_head=nrm_t[:256],
                nn_idx=nn_idx.reshape(-1), nn_dist=nn_dist.reshape(-1))
    I = np.eye(4, dtype=np.float32)
    for name, reg in oracle.REG.items():
        H, b, e, inl = oracle.linearize(reg, oracle.LOSS["HUBER"], src_ds, cov_s, tgt_ds, cov_t, nrm_t, nn_idx,
                                        nn_dist, I, 4.0, 1.0, mode=1)
        gold[f"H_{name}"], gold[f"b_{name}"], gold[f"e_{name}"], gold[f"inl_{name}"] = H, b, e, inl
    for opt_name, opt in oracle.OPT.items():
        P = oracle.default_params(reg_type=3, loss=oracle.LOSS["HUBER"], opt_method=opt, max_iterations=20)
        r = oracle.align(P, src_ds, cov_s, tgt_ds, cov_t, nrm_t, tree_t)
        gold[f"T_GICP_{opt_name}"] = r["T"]
        gold[f"iters_GICP_{opt_name}"] = r["iterations"]
        print(opt_name, r["iterations"], r["converged"], "\n", r["T"])
    np.savez_compressed(os.path.join(OUT, "bundled_pair_golden.npz"), **gold)
    print("T_gt\n", T_gt)


if __name__ == "__main__":
    main()
