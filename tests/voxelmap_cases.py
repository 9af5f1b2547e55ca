"""The known answers of the reference's own tests for mapping::VoxelHashMap (T/test_voxel_hash_map.cpp), restated
against an abstract map interface so that the SAME assertions run on the oracle (CPU, tests/test_oracle_golden.py)
and on the CUDA path through the C-ABI (tests/test_gpu_voxelmap.py).

`make(voxel_size)` returns an adapter with: set(**params), add(points_xyz, pose=None, covs=None, rgb=None,
intensities=None), down(center=(0,0,0), distance=100) -> dict(points, covs, rgb, intensities), overlap(points_xyz, pose)."""
import numpy as np
import scipy.linalg as sl


def xyz1(a):
    a = np.asarray(a, np.float32).reshape(-1, 3)
    return np.c_[a, np.ones(len(a), np.float32)].astype(np.float32)


def make_cov(xx, xy, xz, yy, yz, zz):  # T/test_voxel_hash_map.cpp:61-73
    c = np.zeros((4, 4), np.float32)
    c[0, 0], c[0, 1], c[0, 2], c[1, 1], c[1, 2], c[2, 2] = xx, xy, xz, yy, yz, zz
    c[1, 0], c[2, 0], c[2, 1] = xy, xz, yz
    return c


def expect_cov(covs, pose=None):  # :75-88 (log-Euclidean mean, then rotated), in fp64
    L = sum(sl.logm(c[:3, :3].astype(np.float64)) for c in covs) / len(covs)
    E = sl.expm(L).real
    if pose is not None:
        R = np.asarray(pose, np.float64)[:3, :3]
        E = R @ E @ R.T
    return E


def sort_rows(p):
    return p[np.lexsort((p[:, 2], p[:, 1], p[:, 0]))]


def case_aggregates_points(make):  # :101-147
    m = make(0.1)
    m.add([[0.02, 0.02, 0], [0.03, 0.04, 0], [0.11, 0.02, 0], [0.12, 0.03, 0]])
    p = sort_rows(m.down()["points"])
    assert p.shape == (2, 4)
    np.testing.assert_allclose(p[:, :3], [[0.025, 0.03, 0], [0.115, 0.025, 0]], atol=1e-5)
    assert (p[:, 3] == 1).all()


def case_rgb_intensity(make):  # :149-191
    m = make(0.5)
    m.add([[0, 0, 0], [0.1, 0, 0]], rgb=[[0.2, 0.4, 0.6, 1.0], [0.6, 0.2, 0.0, 1.0]], intensities=[10.0, 20.0])
    r = m.down()
    assert len(r["points"]) == 1 and r["rgb"] is not None and r["intensities"] is not None and r["covs"] is None
    np.testing.assert_allclose(r["points"][0, :3], [0.05, 0, 0], atol=1e-5)
    np.testing.assert_allclose(r["rgb"][0], [0.4, 0.3, 0.3, 1.0], atol=1e-5)
    np.testing.assert_allclose(r["intensities"][0], 15.0, atol=1e-5)


def case_covariances(make):  # :193-247
    covs = [make_cov(1.0, 0.2, 0.3, 2.0, 0.4, 3.0), make_cov(3.0, 0.6, 0.9, 4.0, 0.8, 5.0)]
    m = make(0.5)
    m.add([[0, 0, 0], [0.1, 0, 0]], covs=covs, rgb=[[0.2, 0.4, 0.6, 1.0], [0.6, 0.2, 0.0, 1.0]], intensities=[10.0, 20.0])
    r = m.down()
    assert len(r["points"]) == 1 and r["covs"] is not None
    c = r["covs"][0]
    np.testing.assert_allclose(c[:3, :3], expect_cov(covs), atol=1e-5)
    assert np.abs(c[3, :]).max() == 0 and np.abs(c[:, 3]).max() == 0
    np.testing.assert_allclose(r["rgb"][0], [0.4, 0.3, 0.3, 1.0], atol=1e-5)
    np.testing.assert_allclose(r["intensities"][0], 15.0, atol=1e-5)


def case_rotates_covariances(make):  # :249-289
    covs = [make_cov(1, 0, 0, 4, 0, 9), make_cov(9, 0, 0, 16, 0, 25)]
    pose = np.eye(4, dtype=np.float32)
    a = np.float32(np.pi) / np.float32(2)
    pose[:3, :3] = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]], np.float32)
    pose[:3, 3] = [1, 0, 0]
    m = make(0.5)
    m.add([[0, 0, 0], [0.1, 0, 0]], pose=pose, covs=covs)
    r = m.down()
    assert len(r["points"]) == 1
    np.testing.assert_allclose(r["covs"][0][:3, :3], expect_cov(covs, pose), atol=1e-4)


def case_no_cov_without_input(make):  # :291-308
    m = make(0.5)
    m.add([[0, 0, 0], [0.1, 0, 0]])
    r = m.down()
    assert len(r["points"]) == 1 and r["covs"] is None


def case_min_num_point(make):  # :310-338
    m = make(0.2)
    m.set(min_num_point=2)
    m.add([[0.01, 0.01, 0], [0.02, 0.01, 0], [0.30, 0.30, 0]])
    p = m.down()["points"]
    assert len(p) == 1
    np.testing.assert_allclose(p[0, :3], [0.015, 0.01, 0], atol=1e-5)


def case_bounding_box(make):  # :340-370
    m = make(0.2)
    m.add([[1.05, 0, 0], [1.12, 0, 0], [1.35, 0, 0], [1.00, 0.25, 0]])
    p = m.down(center=(1.0, 0.0, 0.0), distance=0.2)["points"]
    assert len(p) == 1
    np.testing.assert_allclose(p[0, :3], [1.085, 0, 0], atol=1e-5)


def case_overlap_ratio(make):  # :372-410
    m = make(0.5)
    map_pts = [[0.1, 0.1, 0], [1.1, 0, 0]]
    m.add(map_pts)
    query = [[-0.9, 0.1, 0], [0.1, 0, 0], [1.0, 0, 0]]
    pose = np.eye(4, dtype=np.float32)
    pose[0, 3] = 1.0
    assert abs(m.overlap(query, pose) - 2.0 / 3.0) < 1e-5
    m.set(min_num_point=2)
    assert abs(m.overlap(query, pose)) < 1e-5
    m.add(map_pts)
    assert abs(m.overlap(query, pose) - 2.0 / 3.0) < 1e-5


def case_large_batch(make):  # :412-440
    m = make(1.0)
    x = np.arange(100, dtype=np.float32) * 2.0 + 0.5
    m.add(np.c_[x, np.full(100, 0.5), np.full(100, 0.5)])
    p = m.down(distance=1000.0)["points"]
    assert len(p) == 100
    np.testing.assert_allclose(np.sort(p[:, 0]), x, atol=1e-5)


def case_rehash(make):  # :442-483
    m = make(1.0)
    m.set(rehash_threshold=0.0)
    m.add([[0.5, 0.5, 0.5], [10.5, 0.5, 0.5], [20.5, 0.5, 0.5]])
    cap0 = m.info()["capacity"]
    m.add([[30.5, 0.5, 0.5], [40.5, 0.5, 0.5]])
    assert m.info()["capacity"] > cap0  # voxel_num / capacity > 0 -> the next prime
    p = m.down()["points"]
    assert len(p) == 5
    np.testing.assert_allclose(np.sort(p[:, 0]), [0.5, 10.5, 20.5, 30.5, 40.5], atol=1e-5)


def case_staleness(make):  # :485-520
    m = make(0.1)
    m.set(max_staleness=1, remove_old_data_cycle=1)
    m.add([[0, 0, 0]])
    assert len(m.down()["points"]) == 1
    m.add([[1.0, 0, 0]])
    assert len(m.down()["points"]) == 2
    m.add(np.zeros((0, 3), np.float32))
    p = m.down()["points"]
    assert len(p) == 1
    np.testing.assert_allclose(p[0, :3], [1.0, 0, 0], atol=1e-5)


CASES = [case_aggregates_points, case_rgb_intensity, case_covariances, case_rotates_covariances,
         case_no_cov_without_input, case_min_num_point, case_bounding_box, case_overlap_ratio, case_large_batch,
         case_rehash, case_staleness]
