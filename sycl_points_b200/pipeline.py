"""pipeline::lidar_odometry — the production caller of the registration hot path, as a
host-side mirror of the reference's classes over the same C-ABI kernels:

    PCProcessor            I/pipeline/pointcloud_processing.hpp:29-205   box filter -> polar grid -> voxel grid ->
                                                                       random sampling; covariances; refine filter
    Submap                 I/pipeline/submapping.hpp:20-249              keyframe decision, VoxelHashMap submap, target
                                                                       index + covariances / normals for the factor
    AdaptiveMotionPredictor / MotionPredictor
                           I/pipeline/adaptive_motion_predictor.hpp:17-143, motion_predictor.hpp:50-84
    LiDAROdometryPipeline  I/pipeline/lidar_odometry.hpp:27-621          process(scan, timestamp)

Every per-point operation runs in libspx (no CPU fallback); only the per-frame control flow and the 3x3 / 4x4
pose algebra live here, as they live on the host in the reference.  Out of scope and refused loudly rather than
approximated: the IMU paths (preintegration, IMU deskew, initial alignment, GYRO_LIDAR_CV / IMU_SE3 prediction — SURVEY
§2 OUT), the OccupancyGridMap submap, and the intensity filters (correction, gaussian, local-mean normalisation)."""
from __future__ import annotations

import math
import time
from dataclasses import dataclass, field

import numpy as np

from . import api
from ._lib import SpxError
from .api import (CoordinateSystem, ExecutionOptions, KDTree, KNNResult, MapPriorParams, OptimizationMethod, PointCloudShared,
                  PolarGrid, PreprocessFilter, RandomSamplingParams, RegistrationParams, RegistrationPipeline,
                  RegistrationPipelineParams, RegistrationResult, RegType, RobustLossType, RobustScheduleParams,
                  VelocityUpdateParams, VoxelGrid, VoxelHashMap, covariance, deskew, transform)
from .api import DegenerateRegularizationParams, Criteria  # noqa: F401

UNSUPPORTED = -3


# ------------------------------------------------------------------ parameters (odometry_common_params.hpp:46-226)
class SubmapMapType:
    OCCUPANCY_GRID_MAP = 0
    VOXEL_HASH_MAP = 1


def SubmapMapType_from_string(s: str) -> int:
    u = s.upper()
    if u == "OCCUPANCY_GRID_MAP":
        return SubmapMapType.OCCUPANCY_GRID_MAP
    if u == "VOXEL_HASH_MAP":
        return SubmapMapType.VOXEL_HASH_MAP
    raise RuntimeError(f"[SubmapMapType_from_string] Invalid submap map type '{s}'")


class MotionPredictionMode:
    LIDAR_CV = 0
    GYRO_LIDAR_CV = 1
    IMU_SE3 = 2


@dataclass
class IntensityCorrectionParams:
    enable: bool = True
    exp: float = 2.0
    scale: float = 1e-3
    min_intensity: float = 0.0
    max_intensity: float = 1.0
    ref_distance: float = 1.0
    angle_exponent: float = 0.0


@dataclass
class VoxelDownsamplingParams:
    enable: bool = False
    size: float = 1.0


@dataclass
class PolarDownsamplingParams:
    enable: bool = True
    distance_size: float = 1.0
    elevation_size: float = 3.0 * math.pi / 180.0
    azimuth_size: float = 3.0 * math.pi / 180.0
    coord_system: str = "CAMERA"


@dataclass
class RandomDownsamplingParams:
    enable: bool = True
    num: int = 5000


@dataclass
class DownsamplingParams:
    voxel: VoxelDownsamplingParams = field(default_factory=VoxelDownsamplingParams)
    polar: PolarDownsamplingParams = field(default_factory=PolarDownsamplingParams)
    random: RandomDownsamplingParams = field(default_factory=RandomDownsamplingParams)


@dataclass
class BoxFilterParams:
    enable: bool = True
    min: float = 2.0
    max: float = 50.0


@dataclass
class AngleIncidenceFilterParams:
    enable: bool = True
    min_angle: float = 0.0
    max_angle: float = 80.0 * math.pi / 180.0


@dataclass
class PreprocessParams:
    box_filter: BoxFilterParams = field(default_factory=BoxFilterParams)
    angle_incidence_filter: AngleIncidenceFilterParams = field(default_factory=AngleIncidenceFilterParams)


@dataclass
class ToggleParams:
    enable: bool = False


@dataclass
class ScanParams:
    intensity_correction: IntensityCorrectionParams = field(default_factory=IntensityCorrectionParams)
    intensity_gaussian: ToggleParams = field(default_factory=ToggleParams)
    intensity_local_mean_norm: ToggleParams = field(default_factory=ToggleParams)
    enhanced_reflectivity: ToggleParams = field(default_factory=ToggleParams)
    downsampling: DownsamplingParams = field(default_factory=DownsamplingParams)
    preprocess: PreprocessParams = field(default_factory=PreprocessParams)


@dataclass
class KeyframeParams:
    inlier_ratio_threshold: float = 0.7
    distance_threshold: float = 2.0
    angle_threshold_degrees: float = 20.0
    time_threshold_seconds: float = 1.0


@dataclass
class SubmapParams:
    map_type: int = SubmapMapType.OCCUPANCY_GRID_MAP  # the reference's default; only VOXEL_HASH_MAP is built here
    voxel_size: float = 1.0
    max_distance_range: float = 30.0
    point_random_sampling_num: int = 512
    weighted_sampling_ratio: float = 0.8
    keyframe: KeyframeParams = field(default_factory=KeyframeParams)


@dataclass
class MEstimationParams:
    enable: bool = True
    type: RobustLossType = RobustLossType.GEMAN_MCCLURE
    mad_scale: float = 1.0
    min_robust_scale: float = 5.0
    max_iterations: int = 1


@dataclass
class CovarianceEstimationParams:
    neighbor_num: int = 10
    m_estimation: MEstimationParams = field(default_factory=MEstimationParams)


@dataclass
class IMUParams:
    enable: bool = False


@dataclass
class RegistrationFactorSection:
    """CommonParameters::Registration (odometry_common_params.hpp:203-206): min_num_points + the factor half of
    RegistrationParams (reg_type, max_correspondence_distance, robust, rotation_constraint, genz)."""
    min_num_points: int = 100
    factor: RegistrationParams = field(default_factory=RegistrationParams)


@dataclass
class AdaptiveAxisParams:
    factor_min: float = 0.2
    factor_max: float = 1.0
    min_eigenvalue_low: float = 1.0
    min_eigenvalue_high: float = 10.0


@dataclass
class AdaptiveParams:
    rotation: AdaptiveAxisParams = field(default_factory=lambda: AdaptiveAxisParams(0.2, 1.0, 5.0, 10.0))
    translation: AdaptiveAxisParams = field(default_factory=AdaptiveAxisParams)


@dataclass
class MotionPredictionParams:  # MotionPredictor::Params (motion_predictor.hpp:54-56) over AdaptiveMotionPredictor::Params
    verbose: bool = False
    velocity_ema_alpha: float = 1.0
    adaptive: AdaptiveParams = field(default_factory=AdaptiveParams)
    mode: int = MotionPredictionMode.GYRO_LIDAR_CV


@dataclass
class LORegistrationParams:  # Parameters::LO::Registration (lidar_odometry_params.hpp:16-24)
    max_iterations: int = 20
    criteria: Criteria = field(default_factory=Criteria)
    optimization: RegistrationParams = field(default_factory=RegistrationParams)  # gn / lm / dogleg / optimization_method
    degenerate_regularization: DegenerateRegularizationParams = field(default_factory=DegenerateRegularizationParams)
    map_prior: MapPriorParams = field(default_factory=MapPriorParams)


@dataclass
class LOPipelineParams:
    robust: RobustScheduleParams = field(default_factory=RobustScheduleParams)
    velocity_update: VelocityUpdateParams = field(default_factory=VelocityUpdateParams)


@dataclass
class LOParams:
    registration: LORegistrationParams = field(default_factory=LORegistrationParams)
    pipeline: LOPipelineParams = field(default_factory=LOPipelineParams)


@dataclass
class Parameters:
    """lidar_odometry::Parameters (lidar_odometry_params.hpp:12-52) over odometry::CommonParameters, same defaults."""
    device: int = 0
    scan: ScanParams = field(default_factory=ScanParams)
    submap: SubmapParams = field(default_factory=SubmapParams)
    covariance_estimation: CovarianceEstimationParams = field(default_factory=CovarianceEstimationParams)
    imu: IMUParams = field(default_factory=IMUParams)
    registration: RegistrationFactorSection = field(default_factory=RegistrationFactorSection)
    registration_sampling: RandomSamplingParams = field(default_factory=RandomSamplingParams)
    initial_pose: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=np.float32))
    motion_prediction: MotionPredictionParams = field(default_factory=MotionPredictionParams)
    lo: LOParams = field(default_factory=LOParams)

    def make_registration_pipeline_params(self) -> RegistrationPipelineParams:  # lidar_odometry_params.hpp:37-50
        f, o = self.registration.factor, self.lo.registration.optimization
        reg = RegistrationParams(reg_type=f.reg_type, max_correspondence_distance=f.max_correspondence_distance,
                                 robust=f.robust, verbose=f.verbose, gn=o.gn, lm=o.lm, dogleg=o.dogleg,
                                 optimization_method=o.optimization_method, genz=f.genz,
                                 rotation_constraint=f.rotation_constraint)
        reg.max_iterations = self.lo.registration.max_iterations
        reg.criteria = self.lo.registration.criteria
        reg.degenerate_reg = self.lo.registration.degenerate_regularization
        reg.map_prior = self.lo.registration.map_prior
        out = RegistrationPipelineParams()
        out.registration = reg
        out.random_sampling = self.registration_sampling
        out.robust = self.lo.pipeline.robust
        out.velocity_update = self.lo.pipeline.velocity_update
        return out


# ------------------------------------------------------------------ small pose algebra (host, fp32 like Eigen's)
def _inv(T: np.ndarray) -> np.ndarray:
    R, t = T[:3, :3], T[:3, 3]
    out = np.eye(4, dtype=np.float32)
    out[:3, :3] = R.T
    out[:3, 3] = -(R.T @ t)
    return out


def _angle_axis(R: np.ndarray):
    """Eigen::AngleAxisf(R): (angle in [0, pi], unit axis) through the library's se3_log (rotation vector)"""
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = R
    w = api.se3_log(T)[:3]
    a = float(np.linalg.norm(w))
    if a < 1e-12:
        return 0.0, np.array([1.0, 0.0, 0.0], np.float32)
    return a, (w / a).astype(np.float32)


def _rot(angle: float, axis: np.ndarray) -> np.ndarray:
    tw = np.zeros(6, np.float32)
    tw[:3] = np.asarray(axis, np.float32) * np.float32(angle)
    return api.se3_exp(tw)[:3, :3]


# ------------------------------------------------------------------ motion prediction
class AdaptiveMotionPredictor:
    """adaptive_motion_predictor.hpp:17-143: constant-velocity prediction, damped along the directions the previous
    registration constrained well (minimum eigenvalue of the rotation / translation block of H_raw per inlier)."""

    def __init__(self, params: MotionPredictionParams):
        self.params = params
        self._lin = None
        self._ang = None

    @staticmethod
    def _factor(block: np.ndarray, inlier: int, ax: AdaptiveAxisParams) -> float:
        ev = np.linalg.eigvalsh(np.asarray(block, np.float64))  # SelfAdjointEigenSolver<Matrix3f> (third-party)
        ratio = float(ev.min()) / float(inlier)
        score = min(max((ratio - ax.min_eigenvalue_low) / max(ax.min_eigenvalue_high - ax.min_eigenvalue_low, 1e-6), 0.0), 1.0)
        return ax.factor_max * (1.0 - score) + ax.factor_min * score

    def predict(self, linear_velocity, angular_velocity, odom, dt, reg_result: RegistrationResult | None, registrated):
        p = self.params
        rot_factor, trans_factor = p.adaptive.rotation.factor_max, p.adaptive.translation.factor_max
        if registrated and reg_result is not None and reg_result.inlier > 0:
            rot_factor = self._factor(reg_result.H_raw[:3, :3], reg_result.inlier, p.adaptive.rotation)
            trans_factor = self._factor(reg_result.H_raw[3:, 3:], reg_result.inlier, p.adaptive.translation)
        a = np.float32(p.velocity_ema_alpha)
        ang_vec = (np.asarray(angular_velocity[1], np.float32) * np.float32(angular_velocity[0])).astype(np.float32)
        lin = np.asarray(linear_velocity, np.float32)
        self._lin = lin if self._lin is None else (a * lin + (np.float32(1) - a) * self._lin).astype(np.float32)
        self._ang = ang_vec if self._ang is None else (a * ang_vec + (np.float32(1) - a) * self._ang).astype(np.float32)
        ang_norm = float(np.linalg.norm(self._ang))
        ang = (ang_norm, self._ang / np.float32(ang_norm)) if ang_norm > 1e-6 else (0.0, np.array([1, 0, 0], np.float32))
        delta_trans = self._lin * np.float32(dt)
        R = odom[:3, :3]
        T = np.eye(4, dtype=np.float32)
        T[:3, 3] = odom[:3, 3] + R @ (delta_trans * np.float32(trans_factor))
        T[:3, :3] = R @ _rot(ang[0] * dt * rot_factor, ang[1])
        return T


class MotionPredictor:
    """motion_predictor.hpp:50-84.  Without an IMU only LIDAR_CV candidates exist: GYRO_LIDAR_CV / IMU_SE3 fall back
    to the LiDAR constant-velocity prediction exactly as the reference does when no candidate is supplied."""

    def __init__(self, params: MotionPredictionParams):
        self.params = params
        self._cv = AdaptiveMotionPredictor(params)

    def predict(self, linear_velocity, angular_velocity, odom, dt, reg_result, registrated):
        return self._cv.predict(linear_velocity, angular_velocity, odom, dt, reg_result, registrated)


# ------------------------------------------------------------------ scan processing
class PCProcessor:
    """pointcloud_processing::PCProcessor (pointcloud_processing.hpp:29-205)."""

    def __init__(self, queue, scan: ScanParams, covs: CovarianceEstimationParams):
        self.queue, self.scan, self.covs = queue, scan, covs
        self.filter = PreprocessFilter(queue)
        d = scan.downsampling
        self.voxel = VoxelGrid(queue, d.voxel.size) if d.voxel.enable else None
        self.polar = None
        if d.polar.enable:
            cs = {"LIDAR": CoordinateSystem.LIDAR, "CAMERA": CoordinateSystem.CAMERA}.get(d.polar.coord_system.upper())
            if cs is None:
                raise RuntimeError(f"[coordinate_system_from_string] Invalid coordinate system '{d.polar.coord_system}'")
            self.polar = PolarGrid(queue, d.polar.distance_size, d.polar.elevation_size, d.polar.azimuth_size, cs)

    def prefilter(self, src: PointCloudShared) -> PointCloudShared:
        """:130-156 — box filter -> polar grid -> voxel grid -> random sampling; the input is never modified"""
        s, cur = self.scan, src
        if s.preprocess.box_filter.enable:
            cur = self.filter.box_filter(cur, s.preprocess.box_filter.min, s.preprocess.box_filter.max,
                                         PointCloudShared(self.queue))
        if self.polar is not None:
            cur = self.polar.downsampling(cur, PointCloudShared(self.queue))
        if self.voxel is not None:
            cur = self.voxel.downsampling(cur, PointCloudShared(self.queue))
        if cur is src:
            cur = self.filter.random_sampling(src, src.size(), PointCloudShared(self.queue))  # dst = src (a copy)
        if s.downsampling.random.enable:
            cur = self.filter.random_sampling(cur, s.downsampling.random.num, PointCloudShared(self.queue))
        return cur

    def prepare_context(self, scan: PointCloudShared):
        return {"tree": KDTree.build(self.queue, scan), "knn": KNNResult()}

    def compute_covariances(self, scan: PointCloudShared, ctx):
        """:158-171"""
        ctx["tree"].knn_search_async(scan, self.covs.neighbor_num, ctx["knn"])
        m = self.covs.m_estimation
        if m.enable:
            covariance.estimate_robust(ctx["knn"], scan, m.type, m.mad_scale, m.min_robust_scale, m.max_iterations)
        else:
            covariance.estimate(ctx["knn"], scan)

    def refine_filter(self, scan: PointCloudShared, ctx) -> PointCloudShared:
        """:173-205 — angle-of-incidence filter; the intensity filters are not built"""
        s = self.scan
        if s.preprocess.angle_incidence_filter.enable:
            a = s.preprocess.angle_incidence_filter
            scan = self.filter.angle_incidence_filter(scan, a.min_angle, a.max_angle, PointCloudShared(self.queue))
        if scan.has_intensity():
            for name, on in (("intensity_correction", s.intensity_correction.enable and not s.enhanced_reflectivity.enable),
                             ("intensity_gaussian", s.intensity_gaussian.enable),
                             ("intensity_local_mean_norm", s.intensity_local_mean_norm.enable)):
                if on:
                    raise SpxError(UNSUPPORTED, f"[PCProcessor::refine_filter] scan.{name} is not built (SURVEY.md §2 OUT); "
                                                "disable it or drop the intensities")
        return scan


# ------------------------------------------------------------------ submap
class Submap:
    """submapping::Submap (submapping.hpp:20-249) with the VOXEL_HASH_MAP map type."""

    def __init__(self, queue, params: Parameters):
        if params.submap.map_type != SubmapMapType.VOXEL_HASH_MAP:
            raise SpxError(UNSUPPORTED, "[Submap] only submap.map_type = VOXEL_HASH_MAP is built (OccupancyGridMap: SURVEY.md §2 OUT)")
        self.queue = queue
        self.p, self.cov_p, self.reg_p = params.submap, params.covariance_estimation, params.registration
        self.last_keyframe_pose = np.asarray(params.initial_pose, np.float32).copy()
        self.last_keyframe_time = -1.0
        self.keyframe_poses = [self.last_keyframe_pose.copy()]
        self.filter = PreprocessFilter(queue)
        self.voxel = VoxelHashMap(queue, self.p.voxel_size)
        self.tree: KDTree | None = None
        self.last_keyframe_pc = PointCloudShared(queue)
        self.submap_pc = PointCloudShared(queue)
        self._knn = KNNResult()

    def get_submap_point_cloud(self):
        return self.submap_pc

    def get_submap_kdtree(self):
        return self.tree

    def add_first_frame(self, cloud, timestamp, current_pose):  # :90-99
        self.last_keyframe_pose = np.asarray(current_pose, np.float32).copy()
        self.keyframe_poses[0] = self.last_keyframe_pose.copy()
        self._build(cloud, self.last_keyframe_pose, True, None)
        self.last_keyframe_time = timestamp

    def add_frame(self, cloud, reg_result: RegistrationResult, inlier_ratio, timestamp, weights=None) -> bool:  # :109-134
        k = self.p.keyframe
        if k.inlier_ratio_threshold > 0.0 and inlier_ratio <= k.inlier_ratio_threshold:
            return False
        if self._is_keyframe(reg_result, timestamp):
            self.last_keyframe_pose = reg_result.T.copy()
            self.last_keyframe_time = timestamp
            self.keyframe_poses.append(reg_result.T.copy())
            self._build(cloud, reg_result.T, False, weights)
            return True
        return False

    def _is_keyframe(self, reg_result, timestamp) -> bool:  # :157-175
        d = _inv(self.last_keyframe_pose) @ reg_result.T
        distance = float(np.linalg.norm(d[:3, 3]))
        angle = abs(_angle_axis(d[:3, :3])[0]) * (180.0 / math.pi)
        dtime = timestamp - self.last_keyframe_time if self.last_keyframe_time > 0.0 else float("inf")
        k = self.p.keyframe
        return distance >= k.distance_threshold or angle >= k.angle_threshold_degrees or dtime >= k.time_threshold_seconds

    def _build(self, cloud, pose, first, weights):  # :177-212
        if weights is not None and len(weights) == cloud.size():
            self.last_keyframe_pc = self.filter.mixed_random_sampling(cloud, weights, self.p.point_random_sampling_num,
                                                                      self.p.weighted_sampling_ratio, PointCloudShared(self.queue))
        else:
            self.last_keyframe_pc = self.filter.random_sampling(cloud, self.p.point_random_sampling_num,
                                                                PointCloudShared(self.queue))
        self.voxel.add_point_cloud(self.last_keyframe_pc, pose)
        tmp = self.voxel.downsampling(None, pose[:3, 3], self.p.max_distance_range)
        if first:
            self.submap_pc = transform.transform_copy(cloud, pose)
        elif tmp.size() >= self.reg_p.min_num_points:
            self.submap_pc = tmp
        self.tree = KDTree.build(self.queue, self.submap_pc)
        self._compute_covariances()

    def _compute_covariances(self):  # :214-247
        rt = self.reg_p.factor.reg_type
        need_cov = rt in (RegType.GICP, RegType.POINT_TO_DISTRIBUTION, RegType.GENZ) or self.reg_p.factor.rotation_constraint.enable
        need_nrm = rt in (RegType.POINT_TO_PLANE, RegType.GENZ)
        has_cov = self.submap_pc.has_cov()
        knn_ready = False

        def ensure():
            nonlocal knn_ready
            if not knn_ready:
                self.tree.knn_search_async(self.submap_pc, self.cov_p.neighbor_num, self._knn)
                knn_ready = True

        if need_nrm:
            ensure()
            if has_cov:
                covariance.extract_normals(self.submap_pc)
            else:
                covariance.estimate_normals(self._knn, self.submap_pc)
        if need_cov and not has_cov:
            ensure()
            covariance.estimate(self._knn, self.submap_pc)


# ------------------------------------------------------------------ the odometry loop
class ResultType:
    success = 0
    first_frame = 1
    waiting_initial_alignment = 2
    error = 100
    old_timestamp = 101
    small_number_of_points = 102


class LiDAROdometryPipeline:
    """lidar_odometry::LiDAROdometryPipeline (lidar_odometry.hpp:27-621), LiDAR-only:

        process(scan, t): prefilter -> covariances -> refine filter -> [first frame: seed the submap]
                          -> motion prediction -> (MAP prior) -> RegistrationPipeline::align against the submap
                          -> keyframe decision + submap update -> velocity / odometry update."""

    _NAMES = ("1. preprocessing", "2. compute covariances", "3. registration", "4. build submap")

    def __init__(self, params: Parameters, queue=None):
        if params.imu.enable:
            raise SpxError(UNSUPPORTED, "[LiDAR Odometry] the IMU paths are not built (SURVEY.md §2 OUT): set imu.enable = false")
        self.params = params
        self.queue = queue if queue is not None else api.DeviceQueue(params.device)
        self.preprocessed_pc = PointCloudShared(self.queue)
        self.is_first_frame = True
        self.odom = np.asarray(params.initial_pose, np.float32).copy()
        self.prev_odom = self.odom.copy()
        self.linear_velocity = np.zeros(3, np.float32)
        self.angular_velocity = (0.0, np.array([1.0, 0.0, 0.0], np.float32))
        self.pc_processor = PCProcessor(self.queue, params.scan, params.covariance_estimation)
        self.submap = Submap(self.queue, params)
        self.registration_pipeline = RegistrationPipeline(self.queue, params.make_registration_pipeline_params())
        self.reg_result = RegistrationResult()
        self.registrated = False
        self.motion_predictor = MotionPredictor(params.motion_prediction)
        self.last_frame_time = -1.0
        self.dt = -1.0
        self.error_message = ""
        self.current_processing_time: dict = {}
        self.total_processing_times = {n: [] for n in self._NAMES}
        self._ctx = None

    # accessors of the reference
    def get_device_queue(self):
        return self.queue

    def get_error_message(self):
        return self.error_message

    def get_odom(self):
        return self.odom

    def get_prev_odom(self):
        return self.prev_odom

    def get_keyframe_poses(self):
        return self.submap.keyframe_poses

    def get_last_keyframe_pose(self):
        return self.submap.last_keyframe_pose

    def get_preprocessed_point_cloud(self):
        return self.preprocessed_pc

    def get_submap_point_cloud(self):
        return self.submap.get_submap_point_cloud()

    def get_registration_result(self):
        return self.reg_result

    def get_current_processing_time(self):
        return self.current_processing_time

    def get_total_processing_times(self):
        return self.total_processing_times

    def _timed(self, name, fn, acc=None):
        t0 = time.perf_counter()
        out = fn()
        self.queue.wait()
        dt = (time.perf_counter() - t0) * 1e3 + (acc or 0.0)
        self.total_processing_times[name].append(dt)
        self.current_processing_time[name] = dt
        return out

    def process(self, scan: PointCloudShared, timestamp: float) -> int:  # :115-298
        p = self.params
        self.error_message = ""
        if self.last_frame_time > 0.0:
            dt = np.float32(timestamp - self.last_frame_time)
            if dt > 0.0:
                self.dt = float(dt)
            else:
                self.error_message = "old timestamp"
                return ResultType.old_timestamp
        self.current_processing_time = {n: 0.0 for n in self._NAMES}
        stage = "preprocess"
        try:
            t0 = time.perf_counter()
            self.preprocessed_pc = self.pc_processor.prefilter(scan)
            self.queue.wait()
            dt_pre = (time.perf_counter() - t0) * 1e3
            stage = "compute_covariances"
            self._timed(self._NAMES[1], self._compute_covariances)
            stage = "refine_filter"
            self.preprocessed_pc = self._timed(
                self._NAMES[0], lambda: self.pc_processor.refine_filter(self.preprocessed_pc, self._ctx), dt_pre)
        except SpxError:
            raise
        except Exception as e:  # the reference reports and returns (:143-176)
            self.error_message = f"{stage}: {e}"
            return ResultType.error
        if self.preprocessed_pc.size() <= p.registration.min_num_points:
            self.error_message = "point cloud size is too small"
            return ResultType.small_number_of_points
        if self.is_first_frame:
            self.submap.add_first_frame(self.preprocessed_pc, timestamp, self.odom)
            self.is_first_frame = False
            self.last_frame_time = timestamp
            return ResultType.first_frame
        self.reg_result = self._timed(self._NAMES[2], self._registration)
        self._timed(self._NAMES[3], lambda: self._submapping(self.reg_result, timestamp))
        if p.lo.pipeline.velocity_update.enable:  # :265-270: the published cloud is deskewed at full resolution
            deskew.deskew_point_cloud_constant_velocity(self.preprocessed_pc, self.preprocessed_pc, self.odom,
                                                        self.reg_result.T, self.dt)
        self.prev_odom = self.odom
        self.odom = self.reg_result.T.copy()
        self.last_frame_time = timestamp
        delta = _inv(self.prev_odom) @ self.odom
        ang, axis = _angle_axis(delta[:3, :3])
        self.linear_velocity = (delta[:3, 3] / np.float32(self.dt)).astype(np.float32)
        self.angular_velocity = (ang / self.dt, axis)
        self.registrated = True
        return ResultType.success

    def _compute_covariances(self):  # :515-530
        p = self.params
        needs = (p.registration.factor.reg_type == RegType.GICP or p.registration.factor.rotation_constraint.enable or
                 p.scan.preprocess.angle_incidence_filter.enable)
        if not needs:
            self._ctx = None
            return
        self._ctx = self.pc_processor.prepare_context(self.preprocessed_pc)
        self.pc_processor.compute_covariances(self.preprocessed_pc, self._ctx)

    def _registration(self) -> RegistrationResult:  # :544-597
        init_T = self.motion_predictor.predict(self.linear_velocity, self.angular_velocity, self.odom, self.dt,
                                               self.reg_result, self.registrated)
        if self.registrated and self.registration_pipeline.registration is not None:
            self.registration_pipeline.registration.set_map_prior_state(self.reg_result, init_T)
        options = ExecutionOptions(dt=self.dt, prev_pose=self.odom.copy())
        return self.registration_pipeline.align(self.preprocessed_pc, self.submap.get_submap_point_cloud(),
                                                self.submap.get_submap_kdtree(), init_T, options)

    def _submapping(self, reg_result: RegistrationResult, timestamp):  # :599-621
        reg_pc = self.registration_pipeline.get_deskewed_point_cloud()
        if reg_pc is None:
            raise RuntimeError("[LiDAR Odometry] get_deskewed_point_cloud() returned nullptr unexpectedly.")
        p = self.params
        weights = None
        if reg_pc.size() > p.submap.point_random_sampling_num:
            scale = p.lo.pipeline.robust.min_scale if p.lo.pipeline.robust.auto_scale else p.registration.factor.robust.default_scale
            weights = self.registration_pipeline.registration.compute_icp_robust_weights(
                reg_pc, self.submap.get_submap_point_cloud(), self.submap.get_submap_kdtree(), reg_result.T, scale)
        n_in = self.registration_pipeline.get_registration_input_point_cloud()
        inlier_ratio = float(reg_result.inlier) / float(n_in.size()) if n_in is not None and n_in.size() > 0 else 0.0
        self.submap.add_frame(reg_pc, reg_result, inlier_ratio, timestamp, weights)
