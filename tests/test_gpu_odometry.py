"""LiDAROdometryPipeline (sycl_points_b200/pipeline.py; reference: pipeline/lidar_odometry.hpp:115-298,544-621,
pointcloud_processing.hpp:62-156, submapping.hpp:163-247) on a synthetic drive: the per-frame state machine, the
trajectory against ground truth, and — frame by frame — the registration it ran against the oracle's align on
exactly the inputs the pipeline used (preprocessed + sampled source with its covariances, the submap cloud with its
covariances, the predicted initial pose)."""
import numpy as np
import pytest

import oracle
import synthetic

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def q(spx):
    return spx.DeviceQueue(0)


def drive(n_frames, step=0.45, yaw_deg=0.4):
    """sensor poses along the street corridor of the synthetic scene + one revolution per pose (sensor frame)"""
    return synthetic.drive(n_frames, step, yaw_deg)


def make_params(pl, spx):
    import bench
    return bench.odometry_params(pl, spx)


def pose_err(Ta, Tb):
    d = np.linalg.inv(Ta.astype(np.float64)) @ Tb.astype(np.float64)
    w = 0.5 * np.array([d[2, 1] - d[1, 2], d[0, 2] - d[2, 0], d[1, 0] - d[0, 1]])
    return float(np.linalg.norm(d[:3, 3])), float(np.arcsin(min(1.0, np.linalg.norm(w))))


def test_odometry_drive_tracks_ground_truth_and_every_registration_matches_oracle(spx, q):
    from sycl_points_b200 import pipeline as pl
    n_frames = 10
    poses, scans = drive(n_frames)
    P = make_params(pl, spx)
    P.initial_pose = poses[0]
    # a fixed number of iterations (criteria off): the submap differs in its last bits from run to run (the voxel map's
    # atomics), so a step norm that lands within 1e-5 of the 1e-3 criterion could otherwise end the GPU's and the
    # oracle's loops one iteration apart
    P.lo.registration.criteria.translation = 0.0
    P.lo.registration.criteria.rotation = 0.0
    P.lo.registration.max_iterations = 6
    pipe = pl.LiDAROdometryPipeline(P, q)
    # record what every registration was given (host copies taken before the submap moves on)
    calls = []
    inner = pipe.registration_pipeline.align

    def recording_align(source, target, target_knn, initial_guess=None, options=None):
        res = inner(source, target, target_knn, initial_guess, options)
        s = pipe.registration_pipeline.get_registration_input_point_cloud()
        calls.append(dict(src=s.points_host(), cov_s=s.covs_host(), tgt=target.points_host(), cov_t=target.covs_host(),
                          T0=np.asarray(initial_guess, np.float32).copy(), res=res))
        return res

    pipe.registration_pipeline.align = recording_align
    codes = []
    for k in range(n_frames):
        codes.append(pipe.process(spx.PointCloudShared(q, scans[k]), 0.1 * k))
        dt, da = pose_err(poses[k], pipe.get_odom())
        assert dt < 0.05 and da < 0.005, f"frame {k}: {dt:.3f} m, {da:.4f} rad off the ground truth"
    assert codes[0] == pl.ResultType.first_frame and all(c == pl.ResultType.success for c in codes[1:])
    assert len(calls) == n_frames - 1
    assert len(pipe.get_keyframe_poses()) >= 4  # 0.45 m per frame against a 0.3 m keyframe distance
    assert pipe.get_submap_point_cloud().size() > 2000 and pipe.get_submap_point_cloud().has_cov()
    assert set(pipe.get_current_processing_time()) == set(pipe._NAMES)
    # the registration of every frame, re-run by the oracle on the same inputs
    OP = oracle.default_params(reg_type=oracle.REG["GICP"], loss=oracle.LOSS["HUBER"], opt_method=0, max_iterations=6,
                               robust_default_scale=1.0, crit_translation=0.0, crit_rotation=0.0)
    for k, c in enumerate(calls):
        assert len(c["src"]) == 3000  # registration_sampling.num of the ~8000 preprocessed points
        o = oracle.align(OP, c["src"], c["cov_s"], c["tgt"], c["cov_t"], None, oracle.KDTree(c["tgt"]), T_init=c["T0"])
        assert c["res"].iterations == o["iterations"], (k, c["res"].iterations, o["iterations"])
        dt, da = pose_err(o["T"], c["res"].T)
        assert dt < 1e-5 and da < 1e-5, f"registration {k}: dt={dt:.2e} da={da:.2e}"
        assert c["res"].inlier == o["inlier"]
    # motion prediction: from the third frame on the initial guess is a constant-velocity step ahead of the last pose
    step0 = np.linalg.norm(calls[0]["T0"][:3, 3] - poses[0][:3, 3])
    assert step0 < 1e-6  # first registration: no velocity yet -> starts from the previous pose
    later = np.linalg.norm(calls[3]["T0"][:3, 3] - calls[2]["res"].T[:3, 3])
    assert 0.05 < later < 0.6  # (adaptive factor in [0.2, 1] of the 0.45 m step)


def test_odometry_state_machine_and_refusals(spx, q):
    from sycl_points_b200 import pipeline as pl
    poses, scans = drive(3)
    P = make_params(pl, spx)
    pipe = pl.LiDAROdometryPipeline(P, q)
    assert pipe.process(spx.PointCloudShared(q, scans[0][:50]), 0.0) == pl.ResultType.small_number_of_points
    assert pipe.get_error_message() == "point cloud size is too small"
    assert pipe.process(spx.PointCloudShared(q, scans[0]), 1.0) == pl.ResultType.first_frame
    assert pipe.process(spx.PointCloudShared(q, scans[1]), 0.5) == pl.ResultType.old_timestamp  # :130-137
    assert pipe.process(spx.PointCloudShared(q, scans[1]), 1.1) == pl.ResultType.success
    # out-of-scope configurations are refused, not approximated
    P2 = make_params(pl, spx)
    P2.imu.enable = True
    with pytest.raises(spx.SpxError):
        pl.LiDAROdometryPipeline(P2, q)
    P3 = pl.Parameters()  # the reference's default submap is the occupancy grid
    with pytest.raises(spx.SpxError):
        pl.LiDAROdometryPipeline(P3, q)


def test_odometry_with_map_prior_and_velocity_update(spx, q):
    """MAP prior armed from the previous result (lidar_odometry.hpp:572-576) + the constant-velocity deskew wrapper
    (velocity_update, registration_pipeline.hpp:99-110) + the robust-scale schedule: the loop still tracks."""
    from sycl_points_b200 import pipeline as pl
    n_frames = 6
    poses, scans = drive(n_frames)
    P = make_params(pl, spx)
    P.initial_pose = poses[0]
    P.lo.registration.map_prior.enabled = True
    P.lo.pipeline.robust.auto_scale = True
    P.lo.pipeline.robust.init_scale = 4.0
    P.lo.pipeline.robust.min_scale = 0.5
    P.lo.pipeline.robust.auto_scaling_iter = 3
    pipe = pl.LiDAROdometryPipeline(P, q)
    for k in range(n_frames):
        assert pipe.process(spx.PointCloudShared(q, scans[k]), 0.1 * k) in (pl.ResultType.first_frame, pl.ResultType.success)
        dt, da = pose_err(poses[k], pipe.get_odom())
        assert dt < 0.06 and da < 0.006, f"frame {k}: {dt:.3f} m, {da:.4f} rad"
