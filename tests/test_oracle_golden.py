"""Pin the oracle (oracle/spx_oracle.cpp) against every known answer the reference's own tests
hold for the hot path.  T/ = /root/reference/cpp/tests/.  CPU only."""
import numpy as np
import pytest

import oracle


def tie_group_equal(idx_a, dist_a, idx_b, dist_b, eps=1e-4):
    """The reference's comparator (T/test_kdtree.cpp:203-275): sort by (dist, idx), group
    distances within eps, require equal index sets per group."""
    for q in range(len(idx_a)):
        a = sorted(zip(dist_a[q].tolist(), idx_a[q].tolist()))
        b = sorted(zip(dist_b[q].tolist(), idx_b[q].tolist()))
        k = len(a)
        j = 0
        while j < k:
            gd = b[j][0]
            e = j
            while e < k and abs(b[e][0] - gd) <= eps:
                e += 1
            ea = j
            while ea < k and abs(a[ea][0] - gd) <= eps:
                ea += 1
            if ea != e or sorted(x[1] for x in a[j:ea]) != sorted(x[1] for x in b[j:e]):
                return False
            j = e
    return True


@pytest.fixture(scope="module")
def kd_fixture():
    # T/test_kdtree.cpp:21-56: seed 1234, 1000 targets then 100 queries, range 10
    g = oracle.Rng(1234)
    tgt = g.uniform_points(1000, 10.0)
    qry = g.uniform_points(100, 10.0)
    return g, tgt, qry


def test_knn_single_point_known_answer():
    # T/test_kdtree.cpp:358-389: target (0,0,0), query (1,1,1) -> idx 0, dist^2 3.0 +- 1e-6
    tgt = np.array([[0, 0, 0, 1]], np.float32)
    qry = np.array([[1, 1, 1, 1]], np.float32)
    for idx, dist in (oracle.knn_bruteforce(qry, tgt, 1), oracle.KDTree(tgt).knn(qry, 1, mode=0),
                      oracle.KDTree(tgt).knn(qry, 1, mode=1)):
        assert idx[0, 0] == 0
        assert abs(dist[0, 0] - 3.0) <= 1e-6


@pytest.mark.parametrize("k", [1, 3, 5, 10, 20])
def test_kdtree_matches_bruteforce(kd_fixture, k):
    # T/test_kdtree.cpp:301-317,392-408
    _, tgt, qry = kd_fixture
    ib, db = oracle.knn_bruteforce(qry, tgt, k)
    tree = oracle.KDTree(tgt)
    ie, de = tree.knn(qry, k, mode=0)
    ir, dr = tree.knn(qry, k, mode=1)
    assert np.array_equal(ib, ie) and np.array_equal(db, de)  # exact contract: bit-exact
    assert tie_group_equal(ir, dr, ib, db)  # reference traversal: the reference's own criterion
    assert (ib >= 0).all() and (ib < 1000).all() and (db >= 0).all()  # BasicKNNSearch :278-298


def test_kdtree_various_sizes(kd_fixture):
    # T/test_kdtree.cpp:320-355 — continues the fixture's generator
    g, _, _ = kd_fixture
    for nt in (10, 100, 500):
        for nq in (5, 20):
            tgt = g.uniform_points(nt, 10.0)
            qry = g.uniform_points(nq, 10.0)
            ib, db = oracle.knn_bruteforce(qry, tgt, 3)
            ie, de = oracle.KDTree(tgt).knn(qry, 3, mode=0)
            ir, dr = oracle.KDTree(tgt).knn(qry, 3, mode=1)
            assert np.array_equal(ib, ie) and np.array_equal(db, de)
            assert tie_group_equal(ir, dr, ib, db)


def test_self_search_finds_itself():
    # T/test_kdtree.cpp:467-476: every point finds itself first with distance 0
    tgt = oracle.Rng(1234).uniform_points(1000, 10.0)
    idx, dist = oracle.KDTree(tgt).knn(tgt, 10, mode=0)
    assert np.array_equal(idx[:, 0], np.arange(1000))
    assert (dist[:, 0] == 0).all()
    assert (np.diff(dist, axis=1) >= 0).all()


def test_k_larger_than_targets_pads():
    # result.hpp:21-27: unfilled slots stay -1 / FLT_MAX
    tgt = oracle.Rng(7).uniform_points(3, 1.0)
    idx, dist = oracle.knn_bruteforce(tgt, tgt, 5)
    assert (idx[:, 3:] == -1).all() and (dist[:, 3:] == np.finfo(np.float32).max).all()
    idx2, dist2 = oracle.KDTree(tgt).knn(tgt, 5, mode=0)
    assert np.array_equal(idx, idx2) and np.array_equal(dist, dist2)


def test_voxel_grid_known_answer():
    # T/test_downsampling_filters.cpp:27-88
    pts = np.array([[0.10, 0, 0, 1], [0.40, 0, 0, 1], [1.10, 0, 0, 1], [1.40, 0, 0, 1], [0.20, 0, 0, 1]], np.float32)
    rgb = np.array([[10, 20, 30, 1], [20, 40, 60, 1], [30, 60, 90, 1], [50, 70, 90, 1], [70, 80, 90, 1]], np.float32)
    inten = np.array([1, 3, 5, 7, 100], np.float32)
    ts = np.array([0, 2, 4, 6, 8], np.float32)
    out, o_rgb, o_int, o_ts = oracle.voxel_downsample_attrs(pts, 1.0, 2, rgb, inten, ts)
    assert len(out) == 2
    first = [i for i in range(2) if abs(out[i, 0] - 0.233333) < 1e-5]
    assert len(first) == 1
    f = first[0]
    assert abs(o_int[f] - 3.0) < 1e-5
    assert abs(o_ts[f] - 3.333333) < 1e-5
    assert abs(o_rgb[f, 0] - 33.333333) < 1e-5 and abs(o_rgb[f, 1] - 46.666667) < 1e-5
    assert abs(o_rgb[f, 2] - 60.0) < 1e-5
    # output order = ascending key (voxel_constants.hpp:55-61) and the point-only overload agrees
    assert out[0, 0] < out[1, 0]
    assert np.array_equal(oracle.voxel_downsample(pts, 1.0, 2), out)
    # min_voxel_count drops sparse voxels (voxel_downsampling.hpp:204)
    assert len(oracle.voxel_downsample(pts, 1.0, 3)) == 1


def test_voxel_key_layout_and_invalid():
    # voxel_constants.hpp:36-62
    inv = 1.0 / 0.25
    key = oracle.voxel_key([0.3, -0.3, 1.0, 1.0], inv)
    x, y, z = key & 0x1FFFFF, (key >> 21) & 0x1FFFFF, (key >> 42) & 0x1FFFFF
    assert (x, y, z) == (1 + (1 << 20), -2 + (1 << 20), 4 + (1 << 20))
    bad = 2**64 - 1
    assert oracle.voxel_key([np.nan, 0, 0, 1], inv) == bad
    assert oracle.voxel_key([np.inf, 0, 0, 1], inv) == bad
    assert oracle.voxel_key([3e5, 0, 0, 1], inv) == bad  # 1.2e6 cells > 2^20
    pts = np.array([[np.nan, 0, 0, 1], [0.1, 0.1, 0.1, 1]], np.float32)
    assert len(oracle.voxel_downsample(pts, 0.25)) == 1  # invalid points dropped :161-167


def test_eigen_decomposition_reference_matrix():
    # T/test_eigen_utils.cpp:615-623
    A = np.array([[2, 1, 0], [1, 2, 1], [0, 1, 2]], np.float32)
    vals, vecs = oracle.eigen3(A)
    assert np.abs(vecs @ np.diag(vals) @ vecs.T - A).max() <= 1e-5
    assert vals[0] <= vals[1] <= vals[2]


def test_small_matrix_math_vs_numpy():
    # T/test_eigen_utils.cpp:12-15 tolerances: determinant 1e-4 (relative to scale), inverse 1e-3
    rng = np.random.default_rng(0)
    for _ in range(1000):
        A = rng.uniform(-10, 10, (3, 3)).astype(np.float32)
        if abs(np.linalg.det(A.astype(np.float64))) < 1.0:
            continue
        inv = oracle.inverse3(A)
        ref = np.linalg.inv(A.astype(np.float64))
        assert np.abs(inv - ref).max() <= 1e-3 * max(1.0, np.abs(ref).max())
    assert np.array_equal(oracle.inverse3(np.zeros((3, 3))), np.zeros((3, 3), np.float32))  # :406-408


def test_eigen_decomposition_random_spd():
    rng = np.random.default_rng(1)
    for _ in range(500):
        B = rng.normal(size=(3, 3))
        A = (B @ B.T + 1e-3 * np.eye(3)).astype(np.float32)
        vals, vecs = oracle.eigen3(A)
        ref = np.linalg.eigvalsh(A.astype(np.float64))
        assert np.abs(vals - ref).max() <= 2e-4 * max(1.0, ref.max())


def test_so3_se3_exp_log_roundtrip():
    # T/test_eigen_utils.cpp:702-720 (Eigen Random() is uniform in [-1, 1])
    rng = np.random.default_rng(2)
    for _ in range(1000):
        om = rng.uniform(-1, 1, 3).astype(np.float32)
        assert np.abs(oracle.so3_log(oracle.so3_exp(om)) - om).max() <= 1e-5
        tw = rng.uniform(-1, 1, 6).astype(np.float32)
        assert np.abs(oracle.se3_log(oracle.se3_exp(tw)) - tw).max() <= 1e-5
    T = oracle.se3_exp(np.zeros(6, np.float32))
    assert np.array_equal(T, np.eye(4, dtype=np.float32))


def _three_point_case():
    # T/test_registration_pipeline.cpp:79-100
    src = np.array([[0, 0, 0, 1], [1, 0, 0, 1], [5, 0, 0, 1]], np.float32)
    tgt = np.array([[0, 0, 0, 1], [1, 0, 0, 1]], np.float32)
    return src, tgt


def test_robust_weights_none_zero_one():
    # T/test_registration_pipeline.cpp:411-436: P2P, NONE, max_corr 1.5 -> weights {1, 1, 0}
    src, tgt = _three_point_case()
    idx, dist = oracle.knn_bruteforce(src, tgt, 1, np.eye(4))
    w = oracle.robust_weights(0, oracle.LOSS["NONE"], src, None, tgt, None, None, idx, dist, np.eye(4), 1.5**2, 10.0)
    assert w.tolist() == [1.0, 1.0, 0.0]


def test_robust_weights_huber_scale():
    # T/test_registration_pipeline.cpp:477-508: residual 3, HUBER scale 1 -> 1/3, scale 2 -> 2/3
    src = np.array([[3, 0, 0, 1]], np.float32)
    tgt = np.array([[0, 0, 0, 1]], np.float32)
    idx, dist = oracle.knn_bruteforce(src, tgt, 1, np.eye(4))
    for s, want in ((1.0, 1 / 3), (2.0, 2 / 3)):
        w = oracle.robust_weights(0, oracle.LOSS["HUBER"], src, None, tgt, None, None, idx, dist, np.eye(4), 100.0, s)
        assert abs(w[0] - want) <= 1e-5


def test_robust_scale_schedule():
    # T/test_registration_pipeline.cpp:360-409
    s = oracle.robust_schedule(6.0, 2.0, 3)
    assert s[0] == 6.0 and abs(s[1] - np.sqrt(12.0)) <= 1e-5 and abs(s[2] - 2.0) <= 1e-5
    r = oracle.robust_schedule(9.0, 3.0, 3)
    assert r[0] == 9.0 and abs(r[1] - np.sqrt(27.0)) <= 1e-5 and abs(r[2] - 3.0) <= 1e-5


def test_robust_kernels_definitions():
    # robust.hpp:56-114 closed forms
    L = oracle.LOSS
    assert oracle.robust_weight(L["NONE"], 5.0, 1.0) == 1.0
    assert oracle.robust_weight(L["HUBER"], 1e-9, 1.0) == 1.0
    assert oracle.robust_weight(L["TUKEY"], 2.0, 1.0) == 0.0
    assert abs(oracle.robust_weight(L["TUKEY"], 0.5, 1.0) - (1 - 0.25) ** 2) < 1e-7
    assert abs(oracle.robust_weight(L["CAUCHY"], 2.0, 1.0) - 0.2) < 1e-7
    assert abs(oracle.robust_weight(L["GEMAN_MCCLURE"], 2.0, 1.0) - 0.04) < 1e-7
    assert abs(oracle.robust_error(L["NONE"], 3.0, 1.0) - 4.5) < 1e-6
    assert abs(oracle.robust_error(L["HUBER"], 3.0, 1.0) - 2.5) < 1e-6
    assert abs(oracle.robust_error(L["TUKEY"], 3.0, 1.0) - 1 / 6) < 1e-6
    assert abs(oracle.robust_error(L["CAUCHY"], 3.0, 1.0) - 0.5 * np.log(10.0)) < 1e-6
    assert abs(oracle.robust_error(L["GEMAN_MCCLURE"], 3.0, 1.0) - 0.45) < 1e-6


def test_p2p_linearize_hand_computed():
    # factor.hpp:130-149 on one correspondence, T = I: r = pt - ps, J = [skew(ps) | -I]
    src = np.array([[1, 2, 3, 1]], np.float32)
    tgt = np.array([[1.5, 2.0, 2.0, 1]], np.float32)
    idx = np.array([0], np.int32)
    dist = np.array([1.25], np.float32)
    H, b, e, inl = oracle.linearize(0, 0, src, None, tgt, None, None, idx, dist, np.eye(4), 4.0, 1.0, mode=0)
    S = np.array([[0, -3, 2], [3, 0, -1], [-2, 1, 0]], np.float64)
    J = np.hstack([S, -np.eye(3)])
    r = np.array([0.5, 0.0, -1.0])
    assert np.allclose(H, J.T @ J, atol=1e-6) and np.allclose(b, J.T @ r, atol=1e-6)
    assert abs(e - 0.5 * 1.25) < 1e-6 and inl == 1
    # gating: registration.hpp:584 skips dist^2 > max^2
    _, _, e0, inl0 = oracle.linearize(0, 0, src, None, tgt, None, None, idx, dist, np.eye(4), 1.0, 1.0, mode=0)
    assert e0 == 0.0 and inl0 == 0


def test_sum_modes_agree(bundled, bundled_golden):
    src, tgt = bundled["source_ds"], bundled["target_ds"]
    nn_idx, nn_dist = bundled_golden["nn_idx"], bundled_golden["nn_dist"]
    H0, b0, e0, i0 = oracle.linearize(0, 1, src, None, tgt, None, None, nn_idx, nn_dist, np.eye(4), 4.0, 1.0, mode=0)
    H1, b1, e1, i1 = oracle.linearize(0, 1, src, None, tgt, None, None, nn_idx, nn_dist, np.eye(4), 4.0, 1.0, mode=1)
    assert i0 == i1
    assert np.abs(H0 - H1).max() <= 1e-4 * np.abs(H1).max()
    assert np.abs(b0 - b1).max() <= 1e-4 * np.abs(b1).max()
    assert abs(e0 - e1) <= 1e-4 * abs(e1)


def test_bundled_pair_goldens_regress(bundled, bundled_golden):
    """The committed goldens are what the oracle produces today (guards silent oracle drift)."""
    src, tgt = bundled["source_ds"], bundled["target_ds"]
    assert bundled["counts"].tolist() == [69792, 69088, 64625, 63985, 6124, 6096]
    tree = oracle.KDTree(tgt)
    nn_idx, nn_dist = tree.knn(src, 1)
    assert np.array_equal(nn_idx.reshape(-1), bundled_golden["nn_idx"])
    assert np.array_equal(nn_dist.reshape(-1), bundled_golden["nn_dist"])
    idx_s, _ = oracle.KDTree(src).knn(src[:256], 10)
    assert np.array_equal(idx_s, bundled_golden["idx_s_head"])


def test_bundled_pair_lands_near_ground_truth(bundled):
    """cpp/data/T_target_source.txt is asserted by no reference test; SURVEY §8(c) uses it as a
    sanity bound: the restated example (GICP, k=10, voxel 0.25) lands within a few cm / 0.3 deg."""
    src, tgt, T_gt = bundled["source_ds"], bundled["target_ds"], bundled["T_target_source"]
    ti = oracle.KDTree(tgt)
    idx_s, _ = oracle.KDTree(src).knn(src, 10)
    idx_t, _ = ti.knn(tgt, 10)
    cs, ct = oracle.covariance(src, idx_s), oracle.covariance(tgt, idx_t)
    # example_registration.cpp:32-45: GICP, LM, GEMAN_MCCLURE, max_iter 10, auto-scale 10 -> 2.5 in 3 levels
    P = oracle.default_params(reg_type=3, loss=4, opt_method=1, max_iterations=10)
    r = oracle.align_robust(P, src, cs, tgt, ct, None, ti, np.eye(4), 10.0, 2.5, 3)
    dT = np.linalg.inv(T_gt.astype(np.float64)) @ r["T"].astype(np.float64)
    ang = np.degrees(np.arccos(np.clip((np.trace(dT[:3, :3]) - 1) / 2, -1, 1)))
    assert np.linalg.norm(dT[:3, 3]) < 0.05 and ang < 0.3


# ------------------------------------------------------------------ mapping::VoxelHashMap
class _OracleMap:
    """adapter of tests/voxelmap_cases.py over the oracle's sequential restatement"""

    def __init__(self, voxel_size):
        self.m = oracle.VoxelHashMap(voxel_size)

    def set(self, **kw):
        self.m.set_params(**kw)

    def add(self, pts, pose=None, covs=None, rgb=None, intensities=None):
        from voxelmap_cases import xyz1
        cv = None if covs is None else np.array([np.asarray(c, np.float32).T.reshape(16) for c in covs])
        self.m.add_point_cloud(xyz1(pts), pose, cv, None if rgb is None else np.asarray(rgb, np.float32), intensities)

    def down(self, center=(0, 0, 0), distance=100.0):
        r = self.m.downsampling(center, distance)
        if r["covs"] is not None:
            r["covs"] = r["covs"].reshape(-1, 4, 4).transpose(0, 2, 1)
        return r

    def overlap(self, pts, pose=None):
        from voxelmap_cases import xyz1
        return self.m.compute_overlap_ratio(xyz1(pts), pose)

    def info(self):
        return self.m.info()


def _voxelmap_cases():
    import voxelmap_cases
    return voxelmap_cases.CASES


@pytest.mark.parametrize("case", _voxelmap_cases(), ids=lambda c: c.__name__)
def test_voxel_hash_map_reference_known_answers(case):
    """T/test_voxel_hash_map.cpp:92-520, every TEST, on the oracle"""
    case(_OracleMap)


def test_voxel_hash_map_rejects_non_positive_voxel_size():  # T/test_voxel_hash_map.cpp:92-99
    for v in (0.0, -0.1):
        with pytest.raises(ValueError):
            oracle.VoxelHashMap(v)


def test_spd_log_exp_roundtrip_and_scipy():
    """eigen_utils.hpp:646-677 vs scipy.linalg.logm / expm"""
    import scipy.linalg as sl
    rng = np.random.default_rng(5)
    for _ in range(50):
        a = rng.normal(size=(3, 3))
        A = (a @ a.T + 0.05 * np.eye(3)).astype(np.float32)
        L = oracle.spd_function(A, True)
        np.testing.assert_allclose(L, sl.logm(A.astype(np.float64)).real, atol=2e-4 * max(1.0, np.abs(L).max()))
        np.testing.assert_allclose(oracle.spd_function(L, False), A, rtol=0, atol=1e-4 * np.abs(A).max())
