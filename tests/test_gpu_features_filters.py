"""GPU parity: covariance (bit-exact), normals (<= 1e-5: correctly rounded transcendentals on both
sides) and voxel grid / box filter (bit-exact) vs the oracle, through the C-ABI."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def q(spx):
    return spx.DeviceQueue(0)


def test_covariance_bit_exact_and_identity_fallback(spx, q, bundled):
    tgt = bundled["target_ds"]
    cloud = spx.PointCloudShared(q, tgt)
    tree = spx.KDTree.build(q, cloud)
    nn = tree.knn_search(cloud, 10)
    spx.covariance.estimate(nn, cloud)
    want = oracle.covariance(tgt, nn.indices_host())
    got = cloud.covs_host()
    # same fp32 operations in the same order (covariance.hpp:16-47): bit-exact
    assert np.array_equal(got, want)
    assert (got[:, 3, :] == 0).all() and (got[:, :, 3] == 0).all()
    # fewer than 4 valid neighbours -> identity (covariance.hpp:36-42)
    few = spx.PointCloudShared(q, tgt[:3])
    nn3 = spx.KDTree.build(q, few).knn_search(few, 5)
    spx.covariance.estimate(nn3, few)
    c = few.covs_host()
    assert np.array_equal(c[:, :3, :3], np.broadcast_to(np.eye(3, dtype=np.float32), (3, 3, 3)))


def test_normals(spx, q, bundled, bundled_golden):
    tgt = bundled["target_ds"]
    cloud = spx.PointCloudShared(q, tgt)
    nn = spx.KDTree.build(q, cloud).knn_search(cloud, 10)
    spx.covariance.estimate_normals(nn, cloud)
    got = cloud.normals_host()
    want = oracle.normals(tgt, nn.indices_host())
    assert (got[:, 3] == 0).all()
    assert np.abs(np.linalg.norm(got[:, :3], axis=1) - 1).max() < 1e-5
    # same fp32 operations in the same order, transcendentals correctly rounded on both sides
    err = np.abs(got - want).max(axis=1)
    assert np.quantile(err, 0.99) < 1e-5 and err.max() < 1e-5, (np.quantile(err, 0.99), err.max())
    assert np.allclose(got[:256], bundled_golden["nrm_t_head"], atol=1e-5)
    # extract_normals from stored covariances gives the same normals (covariance.hpp:467-495)
    spx.covariance.estimate(nn, cloud)
    cloud.normals = None
    spx.covariance.extract_normals(cloud)
    assert np.abs(cloud.normals_host() - got).max() < 1e-6
    empty = spx.PointCloudShared(q, tgt)
    with pytest.raises(spx.SpxInvalidArgument, match="covariances not computed"):
        spx.covariance.extract_normals(empty)


def test_voxel_reference_known_answer(spx, q):
    # T/test_downsampling_filters.cpp:27-88 (points only: attributes are a "next" row)
    pts = np.array([[0.10, 0, 0, 1], [0.40, 0, 0, 1], [1.10, 0, 0, 1], [1.40, 0, 0, 1], [0.20, 0, 0, 1]], np.float32)
    vg = spx.VoxelGrid(q, 1.0)
    vg.set_min_voxel_count(2)
    out = vg.downsampling(spx.PointCloudShared(q, pts)).points_host()
    assert len(out) == 2
    assert abs(out[0, 0] - 0.233333) < 1e-5 and abs(out[1, 0] - 1.25) < 1e-5
    assert np.array_equal(out, oracle.voxel_downsample(pts, 1.0, 2))
    with pytest.raises(ValueError, match="voxel_size must be positive"):  # voxel_downsampling.hpp:23-25
        spx.VoxelGrid(q, 0.0)
    assert vg.downsampling(spx.PointCloudShared(q, np.zeros((0, 4), np.float32))).size() == 0


@pytest.mark.parametrize("voxel,minc", [(0.25, 1), (0.25, 2), (1.0, 1), (0.05, 1)])
def test_voxel_bundled_scan_bit_exact(spx, q, bundled, voxel, minc):
    raw = bundled["source_raw_head"]
    vg = spx.VoxelGrid(q, voxel)
    vg.set_min_voxel_count(minc)
    got = vg.downsampling(spx.PointCloudShared(q, raw)).points_host()
    want = oracle.voxel_downsample(raw, voxel, minc)
    assert got.shape == want.shape and np.array_equal(got, want)
    if minc == 1:
        assert (got[:, 3] == 1).all()


def test_voxel_invalid_points_and_wide_keys(spx, q):
    rng = np.random.default_rng(4)
    pts = np.c_[rng.uniform(-60, 60, (50000, 3)), np.ones(50000)].astype(np.float32)
    pts[::97, 0] = np.nan
    pts[5::101, 1] = np.inf
    pts[7::103, 2] = 5e5  # outside the 21-bit range at 0.25 m
    got = spx.VoxelGrid(q, 0.25).downsampling(spx.PointCloudShared(q, pts)).points_host()
    assert np.array_equal(got, oracle.voxel_downsample(pts, 0.25))
    # extents x resolution beyond 32 key bits -> 64-bit radix path
    wide = np.c_[rng.uniform(-2000, 2000, (40000, 3)), np.ones(40000)].astype(np.float32)
    wide[:20000] = wide[20000:] + np.float32(0.001)  # make sure there are multi-point voxels
    got = spx.VoxelGrid(q, 0.02).downsampling(spx.PointCloudShared(q, wide)).points_host()
    assert np.array_equal(got, oracle.voxel_downsample(wide, 0.02))
    allbad = np.full((10, 4), np.nan, np.float32)
    assert spx.VoxelGrid(q, 0.5).downsampling(spx.PointCloudShared(q, allbad)).size() == 0
    # w != 1: a voxel whose w sum stays below min_voxel_count = 1 is dropped (voxel_downsampling.hpp:204);
    # the kernel ranks every run first and is run again ranking only the kept ones
    light = pts[np.isfinite(pts).all(axis=1)][:30000].copy()
    light[::3, 3] = np.float32(0.25)
    got = spx.VoxelGrid(q, 0.25).downsampling(spx.PointCloudShared(q, light)).points_host()
    want = oracle.voxel_downsample(light, 0.25)
    assert 0 < len(want) < 30000 and np.array_equal(got, want)


def test_voxel_guessed_key_box_hit_miss_and_size_change(spx, q):
    """The key geometry of a call is guessed from the previous call's box on the same queue (checked on
    the device; exact retry when a point falls outside).  Same output whichever box was used."""
    rng = np.random.default_rng(21)
    small = np.c_[rng.uniform(-10, 10, (40000, 3)), np.ones(40000)].astype(np.float32)
    big = np.c_[rng.uniform(-55, 55, (60000, 3)), np.ones(60000)].astype(np.float32)
    big[::211, 0] = np.nan
    want_small, want_big = oracle.voxel_downsample(small, 0.25), oracle.voxel_downsample(big, 0.25)
    vg = spx.VoxelGrid(q, 0.25)
    first = vg.downsampling(spx.PointCloudShared(q, small)).points_host()   # no guess yet (or a stale one)
    again = vg.downsampling(spx.PointCloudShared(q, small)).points_host()   # guess hits
    assert np.array_equal(first, want_small) and np.array_equal(again, want_small)
    assert np.array_equal(vg.downsampling(spx.PointCloudShared(q, big)).points_host(), want_big)      # guess misses
    assert np.array_equal(vg.downsampling(spx.PointCloudShared(q, small)).points_host(), want_small)  # wider box, same keys order
    assert np.array_equal(spx.VoxelGrid(q, 0.5).downsampling(spx.PointCloudShared(q, big)).points_host(),
                          oracle.voxel_downsample(big, 0.5))                                          # other voxel size: exact box
    shifted = small + np.float32([300, -200, 40, 0])                                                  # disjoint from every earlier box
    assert np.array_equal(spx.VoxelGrid(q, 0.5).downsampling(spx.PointCloudShared(q, shifted)).points_host(),
                          oracle.voxel_downsample(shifted, 0.5))
    allbad = np.full((64, 4), np.nan, np.float32)
    assert spx.VoxelGrid(q, 0.5).downsampling(spx.PointCloudShared(q, allbad)).size() == 0            # guessed, nothing valid
    assert np.array_equal(spx.VoxelGrid(q, 0.5).downsampling(spx.PointCloudShared(q, small)).points_host(),
                          oracle.voxel_downsample(small, 0.5))


def test_voxel_idempotent_and_sorted_large(spx, q):
    """Size-independent properties at a BASELINE-like size: output keys strictly ascending, one
    point per voxel, and down-sampling the output again at the same size is a fixed point."""
    rng = np.random.default_rng(8)
    n = 600000
    pts = np.c_[rng.uniform(-60, 60, (n, 2)), rng.normal(0, 0.3, n), np.ones(n)].astype(np.float32)
    vg = spx.VoxelGrid(q, 0.25)
    out = vg.downsampling(spx.PointCloudShared(q, pts))
    h = out.points_host()
    inv = np.float32(1.0) / np.float32(0.25)
    c = np.floor(h[:, :3] * inv).astype(np.int64) + (1 << 20)
    keys = c[:, 0] | (c[:, 1] << 21) | (c[:, 2] << 42)
    cin = np.floor(pts[:, :3] * inv).astype(np.int64) + (1 << 20)
    assert len(h) == len(np.unique(cin[:, 0] | (cin[:, 1] << 21) | (cin[:, 2] << 42)))
    # a centroid can round onto the voxel boundary, so ascending order is checked on the input keys' order
    assert (np.diff(np.unique(keys)) > 0).all()
    again = vg.downsampling(out).points_host()
    assert len(again) <= len(h)
    sub = pts[:50000]
    assert np.array_equal(vg.downsampling(spx.PointCloudShared(q, sub)).points_host(),
                          oracle.voxel_downsample(sub, 0.25))


def test_box_filter(spx, q, bundled):
    raw = bundled["source_raw_head"].copy()
    raw[3, 0] = np.nan
    raw[10, 3] = np.inf
    cloud = spx.PointCloudShared(q, raw)
    spx.PreprocessFilter(q).box_filter(cloud, 0.5, 50.0)
    assert np.array_equal(cloud.points_host(), oracle.box_filter(raw, 0.5, 50.0))


def test_filters_compact_every_attribute(spx, q, bundled):
    """filter_by_flags (common/filter_by_flags.hpp:29-57): box_filter and random_sampling compact every
    attribute the cloud carries, in source order; the keep-all branch of random_sampling deep-copies."""
    raw = bundled["source_raw_head"][:5000].copy()
    n = len(raw)
    rs = np.random.RandomState(1)
    covs = np.zeros((n, 4, 4), np.float32)
    covs[:, :3, :3] = rs.rand(n, 3, 3)
    nrm, rgb = rs.rand(n, 4).astype(np.float32), rs.rand(n, 4).astype(np.float32)
    inten, ts = rs.rand(n).astype(np.float32), rs.rand(n).astype(np.float32)

    def make():
        c = spx.PointCloudShared(q, raw, covs, nrm)
        c.set_rgb(rgb)
        c.set_intensities(inten)
        c.set_timestamp_offsets(ts)
        return c

    keep = np.array([np.isfinite(p).all() and 2.0 <= np.abs(p[:3]).max() <= 30.0 for p in raw])
    pf = spx.PreprocessFilter(q)
    c = make()
    pf.box_filter(c, 2.0, 30.0)
    assert c.size() == keep.sum() and 0 < keep.sum() < n
    assert np.array_equal(c.points_host(), raw[keep]) and np.array_equal(c.covs_host(), covs[keep])
    assert np.array_equal(c.normals_host(), nrm[keep]) and c.has_rgb() and c.has_intensity() and c.has_timestamps()
    assert np.array_equal(c.rgb.download(), rgb[keep]) and np.array_equal(c.intensities.download(), inten[keep])
    assert np.array_equal(c.timestamp_offsets.download(), ts[keep])
    c = make()
    out = pf.random_sampling(c, 700)
    sel = pf._last_indices.download()
    assert out.size() == 700 and (np.diff(sel) > 0).all()
    assert np.array_equal(out.rgb.download(), rgb[sel]) and np.array_equal(out.timestamp_offsets.download(), ts[sel])
    assert np.array_equal(out.covs_host(), covs[sel])
    dst = spx.PointCloudShared(q)
    pf.random_sampling(c, n + 5, output=dst)  # keep-all: a copy, not an alias
    assert dst.points is not c.points and np.array_equal(dst.points_host(), raw)
    spx.transform.transform(dst, oracle.se3_exp(np.array([0, 0, 0.1, 1, 0, 0], np.float32)))
    assert np.array_equal(c.points_host(), raw)


def test_voxel_attributes_vs_oracle(spx, q, bundled):
    """Cloud overload (voxel_downsampling.hpp:220-288): mean RGB, MEDIAN intensity, mean timestamp per
    voxel in the stable order — bit-exact vs the oracle, incl. even / odd run lengths and ties."""
    pts = bundled["source_raw_head"]
    rs = np.random.RandomState(3)
    n = len(pts)
    rgb = rs.uniform(0, 1, (n, 4)).astype(np.float32)
    inten = np.round(rs.uniform(0, 255, n)).astype(np.float32)  # many ties
    ts = rs.uniform(0, 100, n).astype(np.float32)
    for voxel, minc in ((0.5, 1), (1.0, 2), (3.0, 1)):
        cloud = spx.PointCloudShared(q, pts)
        cloud.set_rgb(rgb)
        cloud.set_intensities(inten)
        cloud.set_timestamp_offsets(ts)
        vg = spx.VoxelGrid(q, voxel)
        vg.set_min_voxel_count(minc)
        out = vg.downsampling(cloud)
        o_p, o_rgb, o_int, o_ts = oracle.voxel_downsample_attrs(pts, voxel, minc, rgb, inten, ts)
        assert out.size() == len(o_p)
        assert np.array_equal(out.points_host(), o_p)
        assert np.array_equal(out.rgb.download(), o_rgb)
        assert np.array_equal(out.intensities.download(), o_int)
        assert np.array_equal(out.timestamp_offsets.download(), o_ts)
    # subset of attributes
    cloud = spx.PointCloudShared(q, pts)
    cloud.set_intensities(inten)
    out = spx.VoxelGrid(q, 1.0).downsampling(cloud)
    o_p, _, o_int, _ = oracle.voxel_downsample_attrs(pts, 1.0, 1, None, inten, None)
    assert np.array_equal(out.points_host(), o_p) and np.array_equal(out.intensities.download(), o_int)
    assert not out.has_rgb() and not out.has_timestamps()


def test_voxel_reference_attribute_known_answer(spx, q):
    """T/test_downsampling_filters.cpp:27-88 through the C-ABI."""
    pts = np.array([[0.10, 0, 0, 1], [0.40, 0, 0, 1], [1.10, 0, 0, 1], [1.40, 0, 0, 1], [0.20, 0, 0, 1]], np.float32)
    cloud = spx.PointCloudShared(q, pts)
    cloud.set_rgb(np.array([[10, 20, 30, 1], [20, 40, 60, 1], [30, 60, 90, 1], [50, 70, 90, 1], [70, 80, 90, 1]], np.float32))
    cloud.set_intensities(np.array([1, 3, 5, 7, 100], np.float32))
    cloud.set_timestamp_offsets(np.array([0, 2, 4, 6, 8], np.float32))
    vg = spx.VoxelGrid(q, 1.0)
    vg.set_min_voxel_count(2)
    out = vg.downsampling(cloud)
    assert out.size() == 2 and out.has_rgb() and out.has_intensity() and out.has_timestamps()
    p = out.points_host()
    first = int(np.argmin(np.abs(p[:, 0] - 0.233333)))
    assert abs(p[first, 0] - 0.233333) < 1e-5
    assert abs(out.intensities.download()[first] - 3.0) < 1e-5
    assert abs(out.timestamp_offsets.download()[first] - 3.333333) < 1e-5
    assert np.allclose(out.rgb.download()[first, :3], [33.333333, 46.666667, 60.0], atol=1e-5)


def test_random_sampling_matches_reference_rng_stream(spx, q, bundled):
    """PreprocessFilter::random_sampling: the selected set equals the oracle's (libstdc++ mt19937(1234) +
    uniform_int_distribution<size_t>), the RNG state carries over between calls, order is preserved."""
    pts = bundled["source_ds"]
    n = len(pts)
    cloud = spx.PointCloudShared(q, pts)
    cloud.set_intensities(np.arange(n, dtype=np.float32))
    f = spx.PreprocessFilter(q)
    rng = oracle.Rng(1234)
    for num in (1000, 1000, 17):
        out = f.random_sampling(cloud, num)
        keep = rng.random_sampling_flags(n, num).astype(bool)
        assert out.size() == num
        assert np.array_equal(out.points_host(), pts[keep])
        assert np.array_equal(out.intensities.download(), np.arange(n, dtype=np.float32)[keep])
    assert f.random_sampling(cloud, n + 5) is cloud  # nothing drawn when the request covers the cloud
    f.set_random_seed(99)
    out = f.random_sampling(cloud, 50)
    assert np.array_equal(out.points_host(), pts[oracle.Rng(99).random_sampling_flags(n, 50).astype(bool)])


def test_cloud_transform_bit_exact(spx, q, bundled):
    """transform::transform (transform.hpp:45-104): points, covariances (T C T^T) and normals in one
    in-place kernel, same fma chains as the reference's eigen_utils::multiply -> bit-exact vs the oracle."""
    tgt = bundled["target_ds"]
    cloud = spx.PointCloudShared(q, tgt)
    nn = spx.KDTree.build(q, cloud).knn_search(cloud, 10)
    spx.covariance.estimate(nn, cloud)
    spx.covariance.estimate_normals(nn, cloud)
    covs, nrm = cloud.covs_host(), cloud.normals_host()
    T = oracle.se3_exp(np.array([0.3, -0.2, 0.5, 1.5, -2.0, 0.25], np.float32))
    o_p, o_c, o_n = oracle.transform_cloud(T, tgt, covs, nrm)
    moved = spx.transform.transform_copy(cloud, T)
    assert np.array_equal(moved.points_host(), o_p)
    assert np.array_equal(moved.covs_host(), o_c)
    assert np.array_equal(moved.normals_host(), o_n)
    assert np.array_equal(cloud.points_host(), tgt)  # the copy left the original alone
    spx.transform.transform(cloud, T)  # in place
    assert np.array_equal(cloud.points_host(), o_p) and np.array_equal(cloud.covs_host(), o_c)
    bare = spx.PointCloudShared(q, tgt)
    spx.transform.transform(bare, T)
    assert np.array_equal(bare.points_host(), o_p) and not bare.has_cov()


@pytest.mark.parametrize("priority", [1, -1])
def test_queue_with_priority_runs_the_same_path(spx, priority):
    """spx_queue_create_with_priority: a scheduling hint only — same results on such a queue."""
    qp = spx.DeviceQueue(0, priority=priority)
    rng = np.random.default_rng(5)
    pts = np.c_[rng.uniform(-20, 20, (30000, 3)), np.ones(30000)].astype(np.float32)
    got = spx.VoxelGrid(qp, 0.5).downsampling(spx.PointCloudShared(qp, pts)).points_host()
    assert np.array_equal(got, oracle.voxel_downsample(pts, 0.5))
    qp.close()


def test_points_from_packed_xyz(spx, q):
    rs = np.random.RandomState(2)
    xyz = rs.normal(0, 30, (10001, 3)).astype(np.float32)
    d = spx.DeviceArray.from_host(q, xyz)
    cloud = spx.PointCloudShared(q)
    cloud.adopt_points(spx.DeviceArray(q, (len(xyz), 4), np.float32), 0)
    cloud.set_points_xyz(d, len(xyz))
    got = cloud.points_host()
    assert np.array_equal(got[:, :3], xyz) and (got[:, 3] == 1.0).all()


@pytest.mark.parametrize("loss", ["NONE", "HUBER", "TUKEY", "CAUCHY", "GEMAN_MCCLURE"])
def test_robust_covariance_bit_exact(spx, q, bundled, loss):
    """covariance::estimate_robust (covariance.hpp:97-134,143-250,323-373): M-estimated covariances, same fp32
    operations in the same order as the oracle -> bit-exact, on the bundled cloud (tiny determinants: the inverse is
    the reference's zero matrix) and on a x10 copy (real Mahalanobis weights), k = 10 / 20, 1 and 3 iterations."""
    for scale_pts, mad, floor_ in ((1.0, 1.0, 1.0), (10.0, 1.5, 1e-3)):
        tgt = bundled["target_ds"].copy()
        tgt[:, :3] *= np.float32(scale_pts)
        cloud = spx.PointCloudShared(q, tgt)
        tree = spx.KDTree.build(q, cloud)
        for k, iters in ((10, 1), (20, 3)):
            nn = tree.knn_search(cloud, k)
            spx.covariance.estimate_robust(nn, cloud, spx.RobustLossType[loss], mad, floor_, iters)
            got = cloud.covs_host()
            want = oracle.covariance_robust(tgt, nn.indices_host(), oracle.LOSS[loss], mad, floor_, iters)
            assert np.array_equal(got, want), (loss, k, iters, np.abs(got - want).max())
            if loss != "NONE" and scale_pts > 1:
                assert not np.array_equal(want, oracle.covariance(tgt, nn.indices_host()))  # the weights did something
    with pytest.raises(RuntimeError, match="neighbor K is too large"):
        spx.covariance.estimate_robust(tree.knn_search(cloud, 65), cloud)


def _deskew_inputs(spx, q, bundled, n=20000):
    tgt = bundled["target_ds"][:n]
    cloud = spx.PointCloudShared(q, tgt)
    nn = spx.KDTree.build(q, cloud).knn_search(cloud, 10)
    spx.covariance.estimate(nn, cloud)
    spx.covariance.estimate_normals(nn, cloud)
    rng = np.random.default_rng(21)
    ts = rng.uniform(-10, 120, len(tgt)).astype(np.float32)  # ms; some before 0 and past the 100 ms scan
    ts[::97] = np.nan
    ts[5::101] = np.inf
    cloud.set_timestamp_offsets(ts)
    return tgt, cloud, ts


def test_deskew_constant_velocity_vs_oracle(spx, q, bundled):
    """deskew::deskew_point_cloud_constant_velocity (relative_pose_deskew.hpp:36-178): same fp32 operations as the
    oracle with correctly rounded sin / cos on both sides -> bit-exact points, normals and covariances; elements
    with a non-finite timestamp are copied; in place == out of place."""
    tgt, cloud, ts = _deskew_inputs(spx, q, bundled)
    covs, nrm = cloud.covs_host(), cloud.normals_host()
    prev = oracle.se3_exp(np.array([0.01, -0.02, 0.03, 0.2, 0.1, -0.05], np.float32))
    cur = oracle.se3_exp(np.array([0.03, -0.01, 0.09, 0.9, 0.3, -0.02], np.float32))
    delta = np.eye(4, dtype=np.float32)
    Rt = prev[:3, :3].T
    delta[:3, :3] = Rt @ cur[:3, :3]
    delta[:3, 3] = Rt @ cur[:3, 3] - Rt @ prev[:3, 3]
    twist = spx.api.se3_log(delta)
    assert np.array_equal(twist, oracle.se3_log(delta))
    o_p, o_c, o_n = oracle.deskew_constant_velocity(tgt, ts, twist, 0.1, covs, nrm)
    out = spx.PointCloudShared(q)
    assert spx.deskew.deskew_point_cloud_constant_velocity(cloud, out, prev, cur, 0.1)
    assert np.array_equal(out.points_host(), o_p)
    assert np.array_equal(out.normals_host(), o_n)
    assert np.array_equal(out.covs_host(), o_c)
    bad = ~np.isfinite(ts)
    assert bad.sum() > 100 and np.array_equal(out.points_host()[bad], tgt[bad])
    assert np.array_equal(out.covs_host()[bad], covs[bad])
    late = np.isfinite(ts) & (ts >= 100.0)  # tau clamps to 1: the whole motion
    whole = oracle.transform_points(oracle.se3_exp(twist), tgt[late])
    assert np.allclose(out.points_host()[late], whole, rtol=0, atol=1e-5)
    assert np.array_equal(out.points_host()[ts <= 0.0], tgt[ts <= 0.0])  # tau = 0: identity motion
    assert np.array_equal(cloud.points_host(), tgt)  # the input was left alone
    assert spx.deskew.deskew_point_cloud_constant_velocity(cloud, cloud, prev, cur, 0.1)  # in place
    assert np.array_equal(cloud.points_host(), o_p) and np.array_equal(cloud.covs_host(), o_c)


def test_deskew_prerequisites_and_pure_translation(spx, q):
    """relative_pose_deskew.hpp:50-61 (false without timestamps / duration) and the translation known answer of
    T/test_relative_pose_deskew.cpp: a point sampled at tau moves by tau x the inter-scan translation."""
    pts = np.array([[1, 0, 0, 1], [0, 2, 0, 1], [0, 0, 3, 1]], np.float32)
    cloud, out = spx.PointCloudShared(q, pts), spx.PointCloudShared(q)
    eye = np.eye(4, dtype=np.float32)
    cur = eye.copy()
    cur[:3, 3] = [1.0, -2.0, 0.5]
    assert not spx.deskew.deskew_point_cloud_constant_velocity(cloud, out, eye, cur, 0.1)  # no timestamps
    cloud.set_timestamp_offsets(np.array([0.0, 50.0, 100.0], np.float32))
    assert not spx.deskew.deskew_point_cloud_constant_velocity(cloud, out, eye, cur, -1.0)  # start == end time
    cloud.start_time_ms, cloud.end_time_ms = 0.0, 100.0
    assert spx.deskew.deskew_point_cloud_constant_velocity(cloud, out, eye, cur, -1.0)  # duration from the cloud
    want = pts.copy()
    want[1, :3] += 0.5 * cur[:3, 3]
    want[2, :3] += cur[:3, 3]
    assert np.allclose(out.points_host(), want, rtol=0, atol=1e-6)
    assert out.has_timestamps() and np.array_equal(out.timestamp_offsets.download(), [0.0, 50.0, 100.0])


def test_velocity_update_aligner_and_pipeline(spx, q):
    """T/test_registration_pipeline.cpp:219-330: accessors before align, the most recent deskewed cloud is a copy
    with its own storage, the fallback without timestamps, and the RegistrationPipeline wiring."""
    def cloud_of(n, stamps=True):
        c = spx.PointCloudShared(q, np.c_[np.arange(n), np.zeros((n, 2)), np.ones(n)].astype(np.float32))
        c.set_intensities(np.arange(n, dtype=np.float32))
        if stamps:
            c.set_timestamp_offsets(np.linspace(0, 100, n).astype(np.float32))
        return c
    seen = []

    def aligner(src, tgt, knn, T, options):
        seen.append((src.size(), src.has_timestamps()))
        r = spx.RegistrationResult(T=np.eye(4, dtype=np.float32))
        r.T[0, 3] = 1.0
        r.inlier = src.size()
        return r

    vu = spx.VelocityUpdateAligner(aligner, 2)
    assert vu.get_deskewed_point_cloud() is None
    src = cloud_of(4)
    opt = spx.ExecutionOptions(dt=0.1)
    res = vu.align(src, cloud_of(3), None, np.eye(4, dtype=np.float32), opt)
    d = vu.get_deskewed_point_cloud()
    assert res.inlier == 4 and seen == [(4, True), (4, True)]
    assert d.size() == 4 and d.has_timestamps() and d.points.ptr.value != src.points.ptr.value
    # second level deskews with T = translate x by 1 over dt: the last point (tau = 1) moved by the full metre
    assert np.allclose(d.points_host()[:, 0], np.arange(4) + np.linspace(0, 1, 4), atol=1e-6)
    seen.clear()
    vu.align(cloud_of(4, stamps=False), cloud_of(3), None, None, opt)
    assert seen == [(4, False)] and not vu.get_deskewed_point_cloud().has_timestamps()
    pp = spx.RegistrationPipelineParams()
    pp.velocity_update.enable, pp.velocity_update.iter = True, 1
    pipe = spx.RegistrationPipeline(aligner, pp)
    assert pipe.get_deskewed_point_cloud() is None
    pipe.align(cloud_of(5), cloud_of(3), None, None, opt)
    assert pipe.get_deskewed_point_cloud().size() == 5


@pytest.mark.parametrize("coord", [0, 1])
def test_polar_grid_vs_oracle(spx, q, bundled, coord):
    """filter::PolarGrid (polar_downsampling.hpp): same key (correctly rounded atan2 on both sides), same stable
    order, same running fp32 sums -> bit-exact points and attributes on the raw bundled scan, both frames."""
    import synthetic
    boxes, cyl = synthetic.make_scene(3)
    raw = synthetic.scan(synthetic.ground_truth_pose(), 64, 2048, boxes, cyl, noise_seed=5)[:120000].copy()
    raw[::1000, :3] = 0.0          # the origin: dropped
    raw[1::1000, :2] = 0.0         # on the LIDAR polar axis
    raw[2::1000, 0] = np.nan
    rng = np.random.default_rng(8)
    n = len(raw)
    rgb = rng.uniform(0, 1, (n, 4)).astype(np.float32)
    inten = rng.uniform(0, 255, n).astype(np.float32)
    ts = rng.uniform(0, 100, n).astype(np.float32)
    cloud = spx.PointCloudShared(q, raw)
    cloud.set_rgb(rgb)
    cloud.set_intensities(inten)
    cloud.set_timestamp_offsets(ts)
    for sizes, minc in (((0.5, 0.02, 0.02), 1), ((2.0, 0.05, 0.1), 3)):
        grid = spx.PolarGrid(q, *sizes, coord)
        grid.set_min_voxel_count(minc)
        out = grid.downsampling(cloud)
        o_p, o_rgb, o_it, o_ts = oracle.polar_downsample_attrs(raw, *sizes, coord, minc, rgb, inten, ts)
        assert out.size() == len(o_p) and 0 < len(o_p) < n
        assert np.array_equal(out.points_host(), o_p)
        assert np.array_equal(out.rgb.download(), o_rgb)
        assert np.array_equal(out.intensities.download(), o_it)
        assert np.array_equal(out.timestamp_offsets.download(), o_ts)
    bare = spx.PolarGrid(q, 0.5, 0.02, 0.02, coord).downsampling(spx.PointCloudShared(q, raw))
    assert np.array_equal(bare.points_host(), oracle.polar_downsample_attrs(raw, 0.5, 0.02, 0.02, coord, 1)[0])
    # a voxel grid call afterwards still finds its own (not the polar) key box
    v = spx.VoxelGrid(q, 0.25).downsampling(spx.PointCloudShared(q, raw))
    assert np.array_equal(v.points_host(), oracle.voxel_downsample(raw, 0.25))


def test_polar_grid_reference_known_answer(spx, q):
    # T/test_downsampling_filters.cpp:90-135
    pts = np.array([[1.1, 0, 0, 1], [1.4, 0, 0, 1], [2.1, 0, 0, 1], [2.3, 0, 0, 1], [2.4, 0, 0, 1]], np.float32)
    cloud = spx.PointCloudShared(q, pts)
    cloud.set_intensities(np.array([2, 4, 6, 10, 100], np.float32))
    grid = spx.PolarGrid(q, 1.0, 3.14159265, 3.14159265, spx.CoordinateSystem.LIDAR)
    grid.set_min_voxel_count(2)
    out = grid.downsampling(cloud)
    assert out.size() == 2 and out.has_intensity()
    got = sorted(zip(out.points_host()[:, 0].tolist(), out.intensities.download().tolist()))
    assert abs(got[0][0] - 1.25) < 1e-5 and abs(got[0][1] - 3.0) < 1e-5
    assert abs(got[1][0] - 2.2666667) < 1e-5 and abs(got[1][1] - 10.0) < 1e-5
    with pytest.raises(ValueError):
        spx.PolarGrid(q, 0.0, 1.0, 1.0)


def test_mixed_random_sampling_matches_reference_rng_stream(spx, q, bundled):
    """PreprocessFilter::mixed_random_sampling (mixed_random_sampling_operator.hpp:29-107): weighted reservoir keys
    for floor(num * ratio) points + partial Fisher-Yates for the rest, the same mt19937 stream as the oracle
    (libstdc++ distributions on both sides), order-preserving compaction of every attribute."""
    pts = bundled["source_ds"]
    n = len(pts)
    rs = np.random.default_rng(8)
    w = rs.uniform(0, 1, n).astype(np.float32)
    w[rs.integers(0, n, 200)] = 0.0  # zero weights never enter the weighted part
    cloud = spx.PointCloudShared(q, pts)
    cloud.set_intensities(np.arange(n, dtype=np.float32))
    f, rng = spx.PreprocessFilter(q), oracle.Rng(1234)
    for num, ratio in ((512, 0.8), (512, 0.8), (100, 0.0), (100, 1.0), (333, 0.5)):
        out = f.mixed_random_sampling(cloud, w, num, ratio)
        keep = rng.mixed_random_sampling_flags(w, num, ratio).astype(bool)
        assert out.size() == keep.sum() == num
        assert np.array_equal(out.points_host(), pts[keep])
        assert np.array_equal(out.intensities.download(), np.arange(n, dtype=np.float32)[keep])
    # the uniform sampler shares nothing with it (separate generators in the reference, one per operator)
    assert f.mixed_random_sampling(cloud, w, n + 1, 0.5) is cloud
    with pytest.raises(ValueError):
        f.mixed_random_sampling(cloud, w, 10, 1.5)
    with pytest.raises(ValueError):
        f.mixed_random_sampling(cloud, np.where(np.arange(n) == 3, -1.0, w).astype(np.float32), 10, 0.5)
    with pytest.raises(ValueError):
        f.mixed_random_sampling(cloud, w[:-1], 10, 0.5)


def test_angle_incidence_filter_matches_oracle(spx, q, bundled):
    """PreprocessFilter::angle_incidence_filter (angle_incidence_filter_operator.hpp:23-111) from covariances and
    from stored normals: the kept set equals the oracle's, attributes follow."""
    tgt = bundled["target_ds"].copy()
    tgt[5, 0] = np.nan
    cloud = spx.PointCloudShared(q, tgt)
    nn = spx.KDTree.build(q, cloud).knn_search(cloud, 10)
    spx.covariance.estimate(nn, cloud)
    covs = cloud.covs_host()
    f = spx.PreprocessFilter(q)
    lo, hi = 0.0, np.float32(80.0 * np.pi / 180.0)
    out = f.angle_incidence_filter(cloud, lo, hi, spx.PointCloudShared(q))
    keep = oracle.angle_incidence_flags(tgt, lo, hi, covs=covs).astype(bool)
    assert 0 < keep.sum() < len(tgt) and not keep[5]
    assert np.array_equal(out.points_host(), tgt[keep])
    assert np.array_equal(out.covs_host(), covs[keep])
    spx.covariance.estimate_normals(nn, cloud)
    nrm = cloud.normals_host()
    out2 = f.angle_incidence_filter(cloud, np.float32(0.2), np.float32(1.2), spx.PointCloudShared(q))
    keep2 = oracle.angle_incidence_flags(tgt, np.float32(0.2), np.float32(1.2), normals=nrm).astype(bool)
    assert np.array_equal(out2.points_host(), tgt[keep2])
    assert np.array_equal(out2.normals_host(), nrm[keep2])
    with pytest.raises(ValueError):
        f.angle_incidence_filter(cloud, 0.5, 0.4)
    with pytest.raises(RuntimeError):
        f.angle_incidence_filter(spx.PointCloudShared(q, tgt), 0.0, 1.0)


def test_weighted_sampling_and_fps_match_oracle(spx, q, bundled):
    """weighted_random_sampling (weighted_sampling_operator.hpp:29-96) and farthest_point_sampling
    (farthest_point_sampling_operator.hpp:27-94; ONE cooperative launch here, a kernel + a host max_element per point
    in the reference): the selected sets equal the oracle's, each operator on its own mt19937(1234) stream."""
    pts = bundled["source_ds"]
    n = len(pts)
    cloud = spx.PointCloudShared(q, pts)
    cloud.set_intensities(np.arange(n, dtype=np.float32))
    f = spx.PreprocessFilter(q)
    w = np.random.default_rng(4).uniform(0, 1, n).astype(np.float32)
    w[::7] = 0.0
    rw, rf, ru = oracle.Rng(1234), oracle.Rng(1234), oracle.Rng(1234)
    for num in (700, 700, 33):
        out = f.weighted_random_sampling(cloud, w, num)
        keep = rw.weighted_random_sampling_flags(w, num).astype(bool)
        assert keep.sum() == num and not keep[::7].any()
        assert np.array_equal(out.points_host(), pts[keep])
    for num in (256, 40):
        out = f.farthest_point_sampling(cloud, num)
        keep = rf.farthest_point_sampling_flags(pts, num).astype(bool)
        assert out.size() == keep.sum() == num
        assert np.array_equal(out.points_host(), pts[keep])
        assert np.array_equal(out.intensities.download(), np.arange(n, dtype=np.float32)[keep])
    # the uniform sampler's stream was not touched by the other operators
    assert np.array_equal(f.random_sampling(cloud, 100).points_host(), pts[ru.random_sampling_flags(n, 100).astype(bool)])
    # duplicates: once every distinct point is taken the maximum distance is 0 and nothing new is added
    dup = np.repeat(pts[:5], 4, axis=0)
    dcloud = spx.PointCloudShared(q, dup)
    got = spx.PreprocessFilter(q).farthest_point_sampling(dcloud, 12)
    want = oracle.Rng(1234).farthest_point_sampling_flags(dup, 12).astype(bool)
    assert np.array_equal(got.points_host(), dup[want]) and got.size() == want.sum() <= 12
    with pytest.raises(ValueError):
        f.weighted_random_sampling(cloud, np.zeros(n, np.float32), 10)
    with pytest.raises(ValueError):
        f.weighted_random_sampling(cloud, np.where(np.arange(n) < 5, 1.0, 0.0).astype(np.float32), 10)
