"""GPU parity: KNN through the C-ABI vs the oracle.  Bit-exact indices and distances, ties by
index (north star).  Mirrors T/test_kdtree.cpp."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
FMAX = np.finfo(np.float32).max


@pytest.fixture(scope="module")
def q(spx):
    return spx.DeviceQueue(0)


def gpu_bf(spx, q, qry, tgt, k, T=None):
    r = spx.knn_search_bruteforce(q, spx.PointCloudShared(q, qry), spx.PointCloudShared(q, tgt), k, T)
    return r.indices_host(), r.distances_host()


def gpu_index(spx, q, qry, tgt, k, T=None, cell=0.0):
    tree = spx.KDTree.build(q, spx.PointCloudShared(q, tgt), cell_size=cell)
    r = tree.knn_search(spx.PointCloudShared(q, qry), k, transT=T)
    return r.indices_host(), r.distances_host(), tree


def assert_same(a, b):
    assert np.array_equal(a[0], b[0]), f"indices differ at {np.argwhere(a[0] != b[0])[:5]}"
    assert np.array_equal(a[1], b[1]), "distances differ"


@pytest.mark.parametrize("k", [1, 3, 5, 10, 20])
def test_fixture_distribution(spx, q, k):
    # T/test_kdtree.cpp:301-317,392-408: seed 1234, 1000 targets, 100 queries, range 10
    g = oracle.Rng(1234)
    tgt, qry = g.uniform_points(1000, 10.0), g.uniform_points(100, 10.0)
    want = oracle.knn_bruteforce(qry, tgt, k)
    assert_same(gpu_bf(spx, q, qry, tgt, k), want)
    assert_same(gpu_index(spx, q, qry, tgt, k)[:2], want)


def test_various_sizes(spx, q):
    # T/test_kdtree.cpp:320-355
    g = oracle.Rng(1234)
    for nt in (10, 100, 500):
        for nq in (5, 20):
            tgt, qry = g.uniform_points(nt, 10.0), g.uniform_points(nq, 10.0)
            want = oracle.knn_bruteforce(qry, tgt, 3)
            assert_same(gpu_bf(spx, q, qry, tgt, 3), want)
            assert_same(gpu_index(spx, q, qry, tgt, 3)[:2], want)


def test_single_point_known_answer(spx, q):
    # T/test_kdtree.cpp:358-389
    tgt = np.array([[0, 0, 0, 1]], np.float32)
    qry = np.array([[1, 1, 1, 1]], np.float32)
    for idx, dist in (gpu_bf(spx, q, qry, tgt, 1), gpu_index(spx, q, qry, tgt, 1)[:2]):
        assert idx[0, 0] == 0 and abs(dist[0, 0] - 3.0) <= 1e-6


def test_self_search(spx, q):
    # T/test_kdtree.cpp:467-476
    tgt = oracle.Rng(1234).uniform_points(1000, 10.0)
    idx, dist, _ = gpu_index(spx, q, tgt, tgt, 10)
    assert np.array_equal(idx[:, 0], np.arange(1000)) and (dist[:, 0] == 0).all()
    assert (np.diff(dist, axis=1) >= 0).all()


def test_ties_resolve_to_lowest_index(spx, q):
    rng = np.random.default_rng(5)
    base = rng.integers(-3, 4, (60, 3)).astype(np.float32)  # integer lattice: many exact distance ties
    tgt = np.c_[np.repeat(base, 4, axis=0), np.ones(240, np.float32)].astype(np.float32)  # + exact duplicates
    tgt = tgt[rng.permutation(len(tgt))]
    qry = np.c_[rng.integers(-3, 4, (80, 3)).astype(np.float32), np.ones(80, np.float32)].astype(np.float32)
    for k in (1, 4, 7, 20):
        want = oracle.knn_bruteforce(qry, tgt, k)
        assert_same(gpu_bf(spx, q, qry, tgt, k), want)
        assert_same(gpu_index(spx, q, qry, tgt, k)[:2], want)
        assert_same(gpu_index(spx, q, qry, tgt, k, cell=0.7)[:2], want)


def test_k_larger_than_targets_and_empty(spx, q):
    tgt = oracle.Rng(7).uniform_points(3, 1.0)
    want = oracle.knn_bruteforce(tgt, tgt, 5)
    got = gpu_bf(spx, q, tgt, tgt, 5)
    assert_same(got, want)
    assert (got[0][:, 3:] == -1).all() and (got[1][:, 3:] == FMAX).all()  # result.hpp:21-27
    assert_same(gpu_index(spx, q, tgt, tgt, 5)[:2], want)
    # empty query -> empty result (kdtree.hpp:429-436); empty target -> all unfilled
    tree = spx.KDTree.build(q, spx.PointCloudShared(q, tgt))
    r = tree.knn_search(spx.PointCloudShared(q, np.zeros((0, 4), np.float32)), 3)
    assert r.query_size == 0
    empty = spx.KDTree.build(q, spx.PointCloudShared(q, np.zeros((0, 4), np.float32)))
    r = empty.knn_search(spx.PointCloudShared(q, tgt), 2)
    assert (r.indices_host() == -1).all() and (r.distances_host() == FMAX).all()


def test_k_limit(spx, q):
    tgt = oracle.Rng(7).uniform_points(8, 1.0)
    tree = spx.KDTree.build(q, spx.PointCloudShared(q, tgt))
    with pytest.raises(spx.SpxInvalidArgument, match="too large"):  # kdtree.hpp:221-223
        tree.knn_search(spx.PointCloudShared(q, tgt), 200)


def test_transform_inside_search(spx, q):
    # KNNBase::knn_search_async transforms queries by transT inside the search (kdtree.hpp:470)
    g = oracle.Rng(99)
    tgt, qry = g.uniform_points(5000, 10.0), g.uniform_points(700, 10.0)
    T = oracle.se3_exp(np.array([0.1, -0.2, 0.3, 1.0, -2.0, 0.5], np.float32))
    want = oracle.knn_bruteforce(qry, tgt, 6, T)
    assert_same(gpu_bf(spx, q, qry, tgt, 6, T), want)
    assert_same(gpu_index(spx, q, qry, tgt, 6, T)[:2], want)
    assert_same(oracle.KDTree(tgt).knn(qry, 6, T, mode=0), want)


def test_queries_far_outside_and_nonfinite(spx, q):
    g = oracle.Rng(3)
    tgt = g.uniform_points(3000, 5.0)
    qry = g.uniform_points(64, 5.0)
    qry[:16, :3] *= 100.0  # far outside the grid: ring search gives up -> full-scan fallback
    qry[16:20, :3] += 40.0
    tgt2 = tgt.copy()
    tgt2[5, 0] = np.nan  # non-finite targets are never neighbours
    tgt2[6, 1] = np.inf
    for k in (1, 5):
        want = oracle.knn_bruteforce(qry, tgt2, k)
        assert_same(gpu_bf(spx, q, qry, tgt2, k), want)
        assert_same(gpu_index(spx, q, qry, tgt2, k)[:2], want)
    qn = qry.copy()
    qn[0, 2] = np.nan
    idx, dist, _ = gpu_index(spx, q, qn, tgt2, 3)
    assert (idx[0] == -1).all() and (dist[0] == FMAX).all()


def test_bundled_pair_k10_and_nn(spx, q, bundled, bundled_golden):
    src, tgt = bundled["source_ds"], bundled["target_ds"]
    want = oracle.KDTree(tgt).knn(tgt, 10)
    got = gpu_index(spx, q, tgt, tgt, 10)
    assert_same(got[:2], want)
    assert np.array_equal(got[0][:256], bundled_golden["idx_t_head"])
    nn = gpu_index(spx, q, src, tgt, 1)
    assert np.array_equal(nn[0].reshape(-1), bundled_golden["nn_idx"])
    assert np.array_equal(nn[1].reshape(-1), bundled_golden["nn_dist"])
    info = got[2].info()
    assert info["n_points"] == len(tgt) and 1.0 <= len(tgt) / info["occupied_cells"] <= 40.0


def test_clustered_and_planar_clouds(spx, q):
    rng = np.random.default_rng(11)
    plane = np.c_[rng.uniform(-50, 50, (20000, 2)), np.zeros(20000)]
    blob = rng.normal(0, 0.05, (5000, 3)) + [10, 10, 2]
    line = np.c_[np.linspace(-30, 30, 3000), np.full(3000, 7.0), np.full(3000, 1.0)]
    tgt = np.c_[np.concatenate([plane, blob, line]), np.ones(28000)].astype(np.float32)
    qry = tgt[rng.choice(len(tgt), 3000, replace=False)].copy()
    qry[:, :3] += rng.normal(0, 0.2, (3000, 3)).astype(np.float32)
    for k in (1, 10, 20):
        want = oracle.KDTree(tgt).knn(qry, k)
        assert_same(gpu_index(spx, q, qry, tgt, k)[:2], want)
    assert_same(gpu_bf(spx, q, qry[:500], tgt, 10), oracle.knn_bruteforce(qry[:500], tgt, 10))


def test_large_100k_k10(spx, q):
    # the reference's own timing case (T/test_kdtree.cpp:411-457): 100k x 100k, k = 10, uniform
    g = oracle.Rng(1234)
    tgt, qry = g.uniform_points(100000, 10.0), g.uniform_points(100000, 10.0)
    want = oracle.KDTree(tgt).knn(qry, 10)
    assert_same(gpu_index(spx, q, qry, tgt, 10)[:2], want)
    sub = np.arange(0, 100000, 50)
    got = gpu_bf(spx, q, qry[sub], tgt, 10)
    assert np.array_equal(got[0], want[0][sub]) and np.array_equal(got[1], want[1][sub])
    # 2-queries-per-thread variant of the tile scan (taken when nq is large)
    got = gpu_bf(spx, q, qry, tgt[:4096], 20)
    assert_same(got, oracle.knn_bruteforce(qry, tgt[:4096], 20))


def _self_knn(spx, q, pts, k):
    cloud = spx.PointCloudShared(q, pts)
    tree = spx.KDTree.build(q, cloud)
    r = tree.knn_search(cloud, k)  # the indexed array itself (covariance estimation's call)
    return (r.indices_host(), r.distances_host()), cloud, tree


@pytest.mark.parametrize("k", [2, 5, 10, 20])
def test_self_search_mixed_density_with_nonfinite(spx, q, k):
    """k-NN of the indexed cloud itself on a scan-like cloud: a plane, a wall, a blob far denser than the
    rest, isolated points hundreds of metres out (the far-query kernel), NaN coordinates."""
    rng = np.random.default_rng(31)
    plane = np.c_[rng.uniform(-40, 40, (30000, 2)), rng.normal(0, 0.02, 30000)]
    wall = np.c_[np.full(6000, 12.0), rng.uniform(-20, 20, 6000), rng.uniform(0, 6, 6000)]
    blob = rng.normal(0, 0.03, (3000, 3)) + [5, -7, 1]
    far = rng.uniform(-300, 300, (40, 3))
    pts = np.c_[np.concatenate([plane, wall, blob, far]), np.ones(39040)].astype(np.float32)
    pts[::997, 1] = np.nan
    pts = pts[rng.permutation(len(pts))]
    got, _, _ = _self_knn(spx, q, pts, k)
    want = oracle.knn_bruteforce(pts, pts, k)  # (a KD-tree built over NaN coordinates is not a reference)
    assert_same(got, want)


def test_self_search_after_the_points_moved(spx, q):
    """Same array, new content after build: queries are searched where they are NOW (the index keeps the
    old positions as targets): moved points, points that became finite / non-finite."""
    rng = np.random.default_rng(32)
    pts = np.c_[rng.uniform(-20, 20, (20000, 2)), rng.normal(0, 0.05, 20000), np.ones(20000)].astype(np.float32)
    pts[::500, 0] = np.nan
    (_, _), cloud, tree = _self_knn(spx, q, pts, 10)
    moved = pts.copy()
    moved[1::7, :3] += rng.normal(0, 1.5, (len(moved[1::7]), 3)).astype(np.float32)   # into other cells
    moved[::1000, 0] = rng.uniform(-20, 20, len(moved[::1000])).astype(np.float32)    # half of the NaNs become finite
    moved[3::900, 2] = np.inf                                                          # some become non-finite
    cloud.points.upload(moved)
    r = tree.knn_search(cloud, 10)
    want = oracle.knn_bruteforce(moved, pts, 10)
    assert_same((r.indices_host(), r.distances_host()), want)


def test_hinted_index_build_is_exact_even_with_a_wrong_box(spx, q, bundled):
    """spx_index_build_hinted: the voxel grid's box + cell edges replace the build's measuring pass; searches stay
    bit-exact vs the oracle — also when the box MISSES most of the cloud (cell coordinates clamp into the grid)."""
    import ctypes as C
    raw = bundled["source_raw_head"]
    vg = spx.VoxelGrid(q, 0.25)
    cloud = vg.downsampling(spx.PointCloudShared(q, raw))
    assert cloud.index_hint is not None
    pts = cloud.points_host()
    lo, hi = cloud.index_hint[0], cloud.index_hint[1]
    assert (pts[:, :3] >= lo).all() and (pts[:, :3] <= hi).all()
    tree = spx.KDTree.build(q, cloud)  # hinted
    oi, od = oracle.knn_bruteforce(pts, pts, 10)
    r = tree.knn_search(cloud, 10)
    assert np.array_equal(r.indices_host(), oi) and np.array_equal(r.distances_host(), od)
    qpts = pts[::3].copy()
    qpts[:, :3] += np.float32(0.1)
    o1, d1 = oracle.knn_bruteforce(qpts, pts, 1)
    qc = spx.PointCloudShared(q, qpts)
    r1 = tree.knn_search(qc, 1)
    assert np.array_equal(r1.indices_host(), o1) and np.array_equal(r1.distances_host(), d1)
    # a box around a corner of the cloud only
    mid = np.median(pts[:, :3], axis=0).astype(np.float32)
    lo2, hi2 = (mid - 1.0).astype(np.float32), (mid + 2.0).astype(np.float32)
    h = C.c_void_p()
    spx._lib.check(spx.lib().spx_index_build_hinted(q.handle, cloud.points.ptr, cloud.size(),
                                                    lo2.ctypes.data_as(C.POINTER(C.c_float)),
                                                    hi2.ctypes.data_as(C.POINTER(C.c_float)), 0.4, 0.6, C.byref(h)))
    t2 = spx.KDTree(q)
    t2._h, t2._n = h, cloud.size()
    r2 = t2.knn_search(cloud, 10)
    assert np.array_equal(r2.indices_host(), oi) and np.array_equal(r2.distances_host(), od)
    r3 = t2.knn_search(qc, 1)
    assert np.array_equal(r3.indices_host(), o1) and np.array_equal(r3.distances_host(), d1)
    # the hint does not survive a transform
    spx.transform.transform(cloud, oracle.se3_exp(np.array([0, 0, 0.3, 5, 0, 0], np.float32)))
    assert cloud.index_hint is None


def test_radius_search_matches_masked_bruteforce(spx, q):
    """kdtree.hpp:251-280 / test_kdtree.cpp:514: max_k nearest within the radius, the rest -1 / FLT_MAX."""
    rng = np.random.default_rng(5)
    tgt = np.c_[rng.uniform(-10, 10, (3000, 3)), np.ones(3000)].astype(np.float32)
    qry = np.c_[rng.uniform(-10, 10, (500, 3)), np.ones(500)].astype(np.float32)
    t, qc = spx.PointCloudShared(q, tgt), spx.PointCloudShared(q, qry)
    tree = spx.KDTree.build(q, t)
    for max_k, radius in ((8, 1.5), (20, 0.4), (1, 2.0)):
        res = tree.radius_search(qc, max_k, radius)
        idx, dist = res.indices_host(), res.distances_host()
        d2 = ((qry[:, None, :3].astype(np.float64) - tgt[None, :, :3].astype(np.float64)) ** 2).sum(-1)
        order = np.argsort(d2, axis=1)[:, :max_k]
        ref_d = np.take_along_axis(d2, order, 1)
        inside = ref_d <= np.float64(np.float32(radius) * np.float32(radius)) * (1 - 1e-6)
        border = np.abs(ref_d - radius * radius) <= 1e-5 * radius * radius
        ok = (idx >= 0)
        assert np.array_equal(ok | border, inside | border)
        m = ok & inside
        np.testing.assert_allclose(dist[m], ref_d[m], rtol=1e-5, atol=1e-6)
        assert np.all(idx[~ok] == -1) and np.all(dist[~ok] == np.finfo(np.float32).max)


def test_remove_nodes_by_flags_renumbers_and_stays_exact(spx, q):
    """kdtree.hpp:282-284 / test_kdtree.cpp:459: after the removal the index answers like a fresh one over the
    kept points, in the compacted numbering."""
    rng = np.random.default_rng(6)
    pts = np.c_[rng.uniform(-10, 10, (4000, 3)), np.ones(4000)].astype(np.float32)
    cloud = spx.PointCloudShared(q, pts)
    tree = spx.KDTree.build(q, cloud)
    flags = (rng.random(4000) > 0.35).astype(np.uint8)
    new_index = np.where(flags == 1, np.cumsum(flags) - 1, -1).astype(np.int32)
    tree.remove_nodes_by_flags(spx.DeviceArray.from_host(q, flags), spx.DeviceArray.from_host(q, new_index))
    kept = spx.PointCloudShared(q, pts[flags == 1])
    got = tree.knn_search(kept, 5)
    want = spx.knn_search_bruteforce(q, kept, kept, 5)
    np.testing.assert_array_equal(got.distances_host(), want.distances_host())
    assert np.array_equal(got.indices_host()[:, 0], np.arange(kept.size()))
    assert tree.info()["n_points"] == kept.size()


def test_handles_may_outlive_their_queue(spx):
    """Garbage-collected callers finalise the members of a reference cycle in no particular order, so a device array,
    an index, a registration or a voxel map can be released AFTER its queue: every release path checks that the queue
    still exists (frees synchronously otherwise) and leaves no pending CUDA error behind for the next launch."""
    q2 = spx.DeviceQueue(0)
    pts = np.random.default_rng(0).uniform(-5, 5, (2000, 4)).astype(np.float32)
    pts[:, 3] = 1
    cloud = spx.PointCloudShared(q2, pts)
    tree = spx.KDTree.build(q2, cloud)
    nn = tree.knn_search(cloud, 5)
    reg = spx.Registration(q2)
    vmap = spx.VoxelHashMap(q2, 0.5)
    vmap.add_point_cloud(cloud)
    q2.wait()
    q2.close()          # the queue goes first ...
    q2.close()          # (twice is harmless)
    tree.close()        # ... then everything that was created on it
    reg.close()
    vmap.close()
    cloud.points.free()
    nn.indices.free()
    nn.distances.free()
    # the library is still healthy: a fresh queue computes and no stale error surfaces at its first launch check
    q3 = spx.DeviceQueue(0)
    c3 = spx.PointCloudShared(q3, pts)
    got = spx.KDTree.build(q3, c3).knn_search(c3, 3).indices_host()
    assert (got[:, 0] == np.arange(len(pts))).all()
