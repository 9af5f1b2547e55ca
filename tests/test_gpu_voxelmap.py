"""GPU parity of mapping::VoxelHashMap (spx_voxelmap_*, through the C-ABI): the reference's own known answers
(T/test_voxel_hash_map.cpp) and the sequential oracle on LiDAR-sized clouds.  The accumulation order inside one
add_point_cloud is unspecified on both sides (fp32 atomics in the reference and here), so sums are compared to a
stated tolerance; the SET of voxels, their counts, the slot arithmetic (capacity, rehash, staleness) are exact."""
import numpy as np
import pytest

import oracle
import synthetic
from voxelmap_cases import CASES, xyz1

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def q(spx):
    return spx.DeviceQueue(0)


class _GpuMap:
    def __init__(self, spx, q, voxel_size):
        self.spx, self.q = spx, q
        self.m = spx.VoxelHashMap(q, voxel_size)

    def set(self, **kw):
        for k, v in kw.items():
            getattr(self.m, "set_" + k)(v)

    def _cloud(self, pts, covs=None, rgb=None, intensities=None):
        c = self.spx.PointCloudShared(self.q, xyz1(pts))
        if covs is not None:
            c.set_covs(np.asarray(covs, np.float32))
        if rgb is not None:
            c.set_rgb(np.asarray(rgb, np.float32))
        if intensities is not None:
            c.set_intensities(np.asarray(intensities, np.float32))
        return c

    def add(self, pts, pose=None, covs=None, rgb=None, intensities=None):
        self.m.add_point_cloud(self._cloud(pts, covs, rgb, intensities), pose)

    def down(self, center=(0, 0, 0), distance=100.0):
        r = self.m.downsampling(None, center, distance)
        n = r.size()
        return {"points": r.points_host(), "covs": r.covs_host() if r.covs is not None and n else None,
                "rgb": r.rgb.download(n) if r.rgb is not None else None,
                "intensities": r.intensities.download(n) if r.intensities is not None else None}

    def overlap(self, pts, pose=None):
        return self.m.compute_overlap_ratio(self._cloud(pts), pose)

    def info(self):
        return self.m.info()


@pytest.mark.parametrize("case", CASES, ids=lambda c: c.__name__)
def test_reference_known_answers_on_gpu(spx, q, case):
    case(lambda voxel: _GpuMap(spx, q, voxel))


def test_rejects_non_positive_voxel_size(spx, q):  # T/test_voxel_hash_map.cpp:92-99
    for v in (0.0, -0.1):
        with pytest.raises(ValueError):
            spx.VoxelHashMap(q, v)
    m = spx.VoxelHashMap(q, 0.5)
    with pytest.raises(ValueError):
        m.set_voxel_size(0.0)


def _by_key(keys, *arrays):
    o = np.argsort(keys, kind="stable")
    return (keys[o],) + tuple(None if a is None else a[o] for a in arrays)


def _compare(spx, gm, om, center=(0, 0, 0), distance=1e4, tol=2e-6):
    """every exported voxel: same key set; centroid / attributes within `tol` of the magnitude of the sums"""
    res, keys = gm.downsampling(None, center, distance, return_keys=True)
    want = om.downsampling(center, distance)
    n = res.size()
    assert n == len(want["keys"]), (n, len(want["keys"]))
    gk, gp = _by_key(keys, res.points_host())
    wk, wp = _by_key(want["keys"], want["points"])
    assert np.array_equal(gk, wk)
    scale = np.maximum(np.abs(wp[:, :3]).max(axis=1, keepdims=True), 1.0)
    assert np.abs(gp[:, :3] - wp[:, :3]).max() <= tol * scale.max(), np.abs(gp[:, :3] - wp[:, :3]).max()
    assert (gp[:, 3] == 1).all()
    ginf, winf = gm.info(), om.info()
    for k in ("capacity", "voxel_num", "staleness_counter", "has_cov", "has_rgb", "has_intensity"):
        assert ginf[k] == winf[k], (k, ginf[k], winf[k])
    if want["covs"] is not None:
        gc = _by_key(keys, res.covs.download(n))[1]
        wc = _by_key(want["keys"], want["covs"])[1]
        s = np.abs(wc).max(axis=1, keepdims=True)
        assert (np.abs(gc - wc) <= 5e-5 * s + 1e-9).all(), (np.abs(gc - wc) / (s + 1e-12)).max()
    if want["rgb"] is not None:
        np.testing.assert_allclose(_by_key(keys, res.rgb.download(n))[1], _by_key(want["keys"], want["rgb"])[1], atol=2e-6)
    if want["intensities"] is not None:
        np.testing.assert_allclose(_by_key(keys, res.intensities.download(n))[1],
                                   _by_key(want["keys"], want["intensities"])[1], rtol=2e-6, atol=1e-5)
    return n


def _se3(rng, t=1.0, a=0.05):
    return oracle.se3_exp(np.r_[rng.uniform(-a, a, 3), rng.uniform(-t, t, 3)].astype(np.float32))


def test_lidar_sequence_matches_oracle(spx, q):
    """a submap built from ten scans with covariances, colours and intensities at moving poses: rehashes twice,
    evicts stale voxels, and every exported voxel agrees with the sequential oracle"""
    rng = np.random.default_rng(3)
    gm, om = spx.VoxelHashMap(q, 0.5), oracle.VoxelHashMap(0.5)
    for m in (gm,):
        m.set_max_staleness(4)
        m.set_remove_old_data_cycle(2)
    om.set_params(max_staleness=4, remove_old_data_cycle=2)
    vg = spx.VoxelGrid(q, 0.25)
    pose = np.eye(4, dtype=np.float32)
    caps = set()
    for f in range(10):
        tgt_raw, _, _ = synthetic.kitti_pair(100 + f, sweeps=1, azimuth_steps=1024)
        cloud = vg.downsampling(spx.PointCloudShared(q, tgt_raw))
        tree = spx.KDTree.build(q, cloud)
        spx.covariance.estimate(tree.knn_search(cloud, 10), cloud)
        n = cloud.size()
        rgb = rng.uniform(0, 1, (n, 4)).astype(np.float32)
        inten = rng.uniform(0, 200, n).astype(np.float32)
        cloud.set_rgb(rgb)
        cloud.set_intensities(inten)
        pose = (pose @ _se3(rng, 2.0, 0.03)).astype(np.float32)
        gm.add_point_cloud(cloud, pose)
        om.add_point_cloud(cloud.points_host(), pose, cloud.covs.download(n), rgb, inten)
        caps.add(gm.info()["capacity"])
        if f in (0, 4, 9):
            _compare(spx, gm, om)
    assert len(caps) >= 2  # at least one rehash happened
    n_all = _compare(spx, gm, om)
    n_box = _compare(spx, gm, om, center=pose[:3, 3], distance=15.0)
    assert 0 < n_box < n_all
    # overlap of the last scan with the map: exact integer arithmetic on both sides
    assert gm.compute_overlap_ratio(cloud, pose) == om.compute_overlap_ratio(cloud.points_host(), pose)
    shifted = pose.copy()
    shifted[:3, 3] += [40.0, 0, 0]
    assert gm.compute_overlap_ratio(cloud, shifted) == om.compute_overlap_ratio(cloud.points_host(), shifted)
    gm.set_min_num_point(3)
    om.set_params(min_num_point=3)
    assert gm.compute_overlap_ratio(cloud, pose) == om.compute_overlap_ratio(cloud.points_host(), pose)
    _compare(spx, gm, om)
    gm.clear()
    assert gm.info()["voxel_num"] == 0 and gm.info()["capacity"] == 30029 and gm.downsampling().size() == 0


def test_invalid_points_and_empty_cloud(spx, q):
    gm, om = spx.VoxelHashMap(q, 1.0), oracle.VoxelHashMap(1.0)
    pts = xyz1([[0.5, 0.5, 0.5], [np.nan, 0, 0], [np.inf, 1, 1], [3e6, 0, 0], [0.6, 0.4, 0.5], [-1048576.5, 0, 0]])
    gm.add_point_cloud(spx.PointCloudShared(q, pts))
    om.add_point_cloud(pts)
    assert gm.info()["voxel_num"] == om.info()["voxel_num"] == 1
    _compare(spx, gm, om, distance=1e7)
    gm.add_point_cloud(spx.PointCloudShared(q))  # N == 0: only the bookkeeping advances (:128-139)
    om.add_point_cloud(np.zeros((0, 4), np.float32))
    assert gm.info() == om.info()


def test_dense_table_probe_sequences(spx, q):
    """96 k distinct voxels pushed through the 30 029 -> 60 013 -> 120 011 -> 240 007 slot tables in chunks that take
    the load to 0.8 before each rehash (long probe sequences, never an exhausted one: which point a full table
    drops depends on the insertion order, in the reference too); keys, counts and sums stay exact"""
    rng = np.random.default_rng(11)
    gm, om = spx.VoxelHashMap(q, 1.0), oracle.VoxelHashMap(1.0)
    for chunk in range(12):
        c = rng.integers(-300, 300, (8000, 3)).astype(np.float32) + 0.5
        pts = xyz1(np.repeat(c, 2, axis=0))  # two identical points per voxel: sums exact in any order
        gm.add_point_cloud(spx.PointCloudShared(q, pts))
        om.add_point_cloud(pts)
        assert gm.info() == om.info()
    _compare(spx, gm, om, tol=0.0)
