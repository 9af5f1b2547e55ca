"""Deterministic synthetic LiDAR-shaped scan pairs for the benchmarks and the full-size parity
tests: a ground plane, 40 axis-aligned buildings and 60 poles inside a
120 m square around a street corridor, scanned by a spinning LiDAR (beams uniform in
[-24.8 deg, +2 deg], range noise N(0, 0.02 m), 100 m max range).

A single revolution of such a sensor fills only ~30 k voxels of 0.25 m (the ground rings are far
apart), so each cloud is the motion-compensated accumulation of `sweeps` revolutions taken every
`spacing` metres along the corridor — what a LiDAR-odometry front end hands to scan matching when
it aligns against a local map.  The default (16 sweeps x 64 beams x 2048 azimuth steps, 8.3 m apart,
~2.05 M raw points) gives the 120 k +- 2 % points after the 0.25 m voxel grid that BASELINE.json
config 2 names.  Target cloud at the identity pose, source cloud at
T_gt = trans(0.5, 0.1, -0.02) * Rz(0.7 deg) — the magnitudes of the reference's bundled
cpp/data/T_target_source.txt.  Pure numpy (no file or network access); seeded with numpy's
MT19937 RandomState, so every box generates the same clouds.
"""
from __future__ import annotations

import numpy as np

GROUND_Z = -1.73


def make_scene(seed: int = 42):
    rs = np.random.RandomState(seed)
    boxes = []
    while len(boxes) < 40:
        cx, cy = rs.uniform(-60, 60, 2)
        sx, sy = rs.uniform(5, 30, 2)
        h = rs.uniform(3, 15)
        lo = np.array([cx - sx / 2, cy - sy / 2, GROUND_Z])
        hi = np.array([cx + sx / 2, cy + sy / 2, GROUND_Z + h])
        # keep the street corridor along x (where the sensor drives) free, with a little clearance
        if lo[1] - 6 < 0 < hi[1] + 6:
            continue
        boxes.append((lo, hi))
    cyl = []
    while len(cyl) < 60:
        cx, cy = rs.uniform(-60, 60, 2)
        if abs(cy) < 4:
            continue
        cyl.append((cx, cy, rs.uniform(0.15, 0.4), GROUND_Z, GROUND_Z + rs.uniform(2, 8)))
    return np.array([np.r_[b[0], b[1]] for b in boxes]), np.array(cyl)


def _az_ranges(origin, R, xy_pts, steps, margin=2):
    """Azimuth-step intervals [(i0, i1), ...] (sensor frame) that can see the given world xy points."""
    v = (np.c_[xy_pts, np.zeros(len(xy_pts))] - np.r_[origin[:2], 0.0]) @ R  # world -> sensor (R^T applied)
    ang = np.arctan2(v[:, 1], v[:, 0])
    ref = ang[0]
    rel = np.angle(np.exp(1j * (ang - ref)))  # objects never span >= 180 deg from outside
    a0, a1 = ref + rel.min(), ref + rel.max()
    i0 = int(np.floor(a0 / (2 * np.pi) * steps)) - margin
    i1 = int(np.ceil(a1 / (2 * np.pi) * steps)) + margin + 1
    if i1 - i0 >= steps:
        return [(0, steps)]
    i0 %= steps
    i1 = i0 + (i1 - int(np.floor(a0 / (2 * np.pi) * steps)) + margin)
    if i1 <= steps:
        return [(i0, i1)]
    return [(i0, steps), (0, i1 - steps)]


def scan(pose: np.ndarray, beams: int, azimuth_steps: int, boxes, cyl, noise_seed: int,
         max_range: float = 100.0) -> np.ndarray:
    """Points of one revolution in the SENSOR frame, (n, 4) float32 xyz1, in firing order
    (azimuth-major).  Objects are only tested against the azimuth columns that can see them."""
    rs = np.random.RandomState(noise_seed)
    el = np.deg2rad(np.linspace(-24.8, 2.0, beams))
    az = np.linspace(0.0, 2 * np.pi, azimuth_steps, endpoint=False)
    A, E = np.meshgrid(az, el, indexing="ij")
    dirs_s = np.stack([np.cos(E) * np.cos(A), np.cos(E) * np.sin(A), np.sin(E)], axis=-1)  # (A, B, 3) sensor frame
    R, o = pose[:3, :3].astype(np.float64), pose[:3, 3].astype(np.float64)
    d = dirs_s @ R.T  # world frame
    t = np.full((azimuth_steps, beams), np.inf)
    with np.errstate(divide="ignore", invalid="ignore"):
        tg = (GROUND_Z - o[2]) / d[..., 2]
        tg[(d[..., 2] >= 0) | (tg <= 0)] = np.inf
        t = np.minimum(t, tg)
        for b in boxes:
            corners = np.array([[b[0], b[1]], [b[0], b[4]], [b[3], b[1]], [b[3], b[4]]])
            for i0, i1 in _az_ranges(o, R, corners, azimuth_steps):
                dd = d[i0:i1]
                inv = 1.0 / dd
                t0 = (b[:3] - o) * inv
                t1 = (b[3:] - o) * inv
                tn = np.minimum(t0, t1).max(axis=-1)
                tf = np.maximum(t0, t1).min(axis=-1)
                hit = (tn <= tf) & (tf > 0) & (tn > 0)
                tt = t[i0:i1]
                t[i0:i1] = np.where(hit & (tn < tt), tn, tt)
        for cx, cy, r, z0, z1 in cyl:
            ring = np.array([[cx - r, cy - r], [cx - r, cy + r], [cx + r, cy - r], [cx + r, cy + r]])
            for i0, i1 in _az_ranges(o, R, ring, azimuth_steps):
                dd = d[i0:i1]
                a = dd[..., 0] ** 2 + dd[..., 1] ** 2
                ox, oy = o[0] - cx, o[1] - cy
                bq = 2 * (ox * dd[..., 0] + oy * dd[..., 1])
                cq = ox * ox + oy * oy - r * r
                disc = bq * bq - 4 * a * cq
                ok = disc > 0
                tc = (-bq - np.sqrt(np.where(ok, disc, 0.0))) / (2 * a)
                z = o[2] + tc * dd[..., 2]
                hit = ok & (tc > 0) & (z >= z0) & (z <= z1)
                tt = t[i0:i1]
                t[i0:i1] = np.where(hit & (tc < tt), tc, tt)
    t[t > max_range] = np.inf
    t = t + rs.normal(0.0, 0.02, t.shape)
    keep = np.isfinite(t)
    p = (dirs_s[keep] * t[keep][:, None]).astype(np.float32)
    return np.concatenate([p, np.ones((len(p), 1), np.float32)], axis=1)


def ground_truth_pose() -> np.ndarray:
    a = np.deg2rad(0.7)
    T = np.eye(4)
    T[:3, :3] = [[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]]
    T[:3, 3] = [0.5, 0.1, -0.02]
    return T


def random_pose(rs: np.random.RandomState, max_t: float = 1.0, max_deg: float = 2.0) -> np.ndarray:
    axis = rs.normal(size=3)
    axis /= np.linalg.norm(axis)
    ang = np.deg2rad(rs.uniform(0, max_deg))
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    T = np.eye(4)
    T[:3, :3] = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
    t = rs.normal(size=3)
    T[:3, 3] = t / np.linalg.norm(t) * rs.uniform(0, max_t)
    T[2, 3] *= 0.1  # vehicles barely move vertically
    return T


def accumulated_cloud(pose: np.ndarray, boxes, cyl, sweeps: int, beams: int, azimuth_steps: int, spacing: float,
                      noise_seed: int) -> np.ndarray:
    """`sweeps` revolutions taken every `spacing` m along the cloud frame's x axis, expressed in the
    cloud frame (whose world pose is `pose`).  (n, 4) float32 xyz1, sweep-major firing order."""
    out = []
    for s in range(sweeps):
        P = np.eye(4)
        P[0, 3] = (s - (sweeps - 1) / 2) * spacing
        p = scan(pose @ P, beams, azimuth_steps, boxes, cyl, noise_seed * 1000 + s)
        q = (p.astype(np.float64) @ P.T).astype(np.float32)
        q[:, 3] = 1.0
        out.append(q)
    return np.concatenate(out)


def kitti_pair(seed: int = 42, sweeps: int = 16, beams: int = 64, azimuth_steps: int = 2048, spacing: float = 8.3,
               T_gt: np.ndarray | None = None):
    """(target_raw, source_raw, T_gt): align(source -> target) should recover T_gt."""
    boxes, cyl = make_scene(seed)
    T_gt = ground_truth_pose() if T_gt is None else np.asarray(T_gt, np.float64)
    tgt = accumulated_cloud(np.eye(4), boxes, cyl, sweeps, beams, azimuth_steps, spacing, seed * 2 + 1)
    src = accumulated_cloud(T_gt, boxes, cyl, sweeps, beams, azimuth_steps, spacing, seed * 2 + 2)
    return tgt, src, T_gt.astype(np.float32)


def drive(n_frames: int, step: float = 0.45, yaw_deg: float = 0.4, seed: int = 42, beams: int = 64,
          azimuth_steps: int = 1024, start_x: float = -20.0):
    """A vehicle driving along the street corridor of scene `seed`: (poses, scans) — the sensor pose of every frame
    (4x4 float32, world frame) and ONE revolution per frame in the sensor frame (~64 k points at 64 x 1024).  The
    input of the odometry loop (bench.py extras, tools/run_odometry.py, tests/test_gpu_odometry.py)."""
    boxes, cyl = make_scene(seed)
    poses, scans = [], []
    T = np.eye(4)
    T[0, 3] = start_x
    a = np.deg2rad(yaw_deg)
    D = np.eye(4)
    D[:3, :3] = [[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]]
    D[0, 3] = step
    for k in range(n_frames):
        poses.append(T.astype(np.float32))
        scans.append(scan(T, beams, azimuth_steps, boxes, cyl, 900 + k))
        T = T @ D
    return poses, scans


def dense_pair(seed: int = 42, n_points: int = 2_000_000):
    """BASELINE config 4 shape: a dense scan pair (8 sweeps x 256 beams x 4096 azimuth steps, ~8.2 M raw
    points per cloud) meant for a 0.05 m voxel grid, which leaves ~1.9 M points per cloud; callers trim
    to `n_points`.  (target_raw, source_raw, T_gt)."""
    return kitti_pair(seed, sweeps=8, beams=256, azimuth_steps=4096, spacing=8.0)


def mt19937_uniform_box(n: int, seed: int, lo, hi) -> np.ndarray:
    """(n, 4) float32 xyz1 with every coordinate drawn as C++'s
    `std::uniform_real_distribution<float>(lo[a], hi[a])(std::mt19937(seed))`, x then y then z per
    point (queries seed 1234, targets seed 4321).  numpy's legacy
    RandomState(seed) is the same init_genrand-seeded MT19937 and its full-range uint32 draws are the
    raw 32-bit outputs; libstdc++'s generate_canonical<float, 24> uses one output per value:
    u = float(x) / 2^32, clamped below 1."""
    rs = np.random.RandomState(seed)
    raw = rs.randint(0, 2**32, size=3 * n, dtype=np.uint64).astype(np.uint32)
    u = raw.astype(np.float32) / np.float32(4294967296.0)
    u = np.minimum(u, np.nextafter(np.float32(1.0), np.float32(0.0)))
    u = u.reshape(n, 3)
    lo, hi = np.asarray(lo, np.float32), np.asarray(hi, np.float32)
    p = np.empty((n, 4), np.float32)
    p[:, :3] = (hi - lo) * u + lo
    p[:, 3] = 1.0
    return p


def knn_config3(n_queries: int = 1_000_000, n_targets: int = 1_000_000):
    """BASELINE config 3 clouds: i.i.d. uniform in [-50, 50]^2 x [-3, 10], mt19937(1234) / (4321)."""
    lo, hi = (-50.0, -50.0, -3.0), (50.0, 50.0, 10.0)
    return mt19937_uniform_box(n_queries, 1234, lo, hi), mt19937_uniform_box(n_targets, 4321, lo, hi)
