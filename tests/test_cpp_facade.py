"""The C++ facade (include/sycl_points/**, header-only over the C-ABI): it must compile with plain
g++ against libspx.so — no SYCL, no Eigen, no CUDA headers — fail loudly without a GPU, and on a
GPU pass the reference's known-answer tests restated in tests/cpp/test_facade.cpp and run the
restated example_registration on the bundled (voxelised) scan pair."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "sycl_points_b200")
BUILD = os.path.join(ROOT, "tests", "cpp", "_build")
GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")


def compile_cpp(src, out):
    import sycl_points_b200  # noqa: F401  (builds libspx.so if missing)
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, out)
    if os.path.exists(exe) and os.path.getmtime(exe) > max(
            os.path.getmtime(os.path.join(ROOT, src)), os.path.getmtime(os.path.join(LIBDIR, "libspx.so")),
            *[os.path.getmtime(os.path.join(dp, f)) for dp, _, fs in os.walk(os.path.join(ROOT, "include")) for f in fs]):
        return exe
    cmd = [GXX, "-std=c++20", "-O1", "-Wall", "-Wextra", "-Werror=return-type", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, src), "-L" + LIBDIR, "-lspx", "-Wl,-rpath," + LIBDIR, "-o", exe]
    env = {k: v for k, v in os.environ.items() if k not in ("CXX", "CC")}
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr[-4000:]
    return exe


@pytest.fixture(scope="module")
def facade_test_exe():
    return compile_cpp("tests/cpp/test_facade.cpp", "test_facade")


@pytest.fixture(scope="module")
def example_exe():
    return compile_cpp("examples/example_registration.cpp", "example_registration")


def test_facade_compiles_without_sycl_or_eigen(facade_test_exe, example_exe):
    assert os.access(facade_test_exe, os.X_OK) and os.access(example_exe, os.X_OK)
    # the facade depends on nothing but the C-ABI: no CUDA / SYCL / Eigen include anywhere in it
    for dp, _, fs in os.walk(os.path.join(ROOT, "include", "sycl_points")):
        for f in fs:
            text = open(os.path.join(dp, f)).read()
            assert "cuda_runtime" not in text and "<sycl/" not in text and "<CL/" not in text, f


def test_facade_fails_loudly_without_gpu(facade_test_exe):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([facade_test_exe], capture_output=True, text=True)
    assert r.returncode != 0
    assert "DeviceQueue" in r.stderr or "CUDA" in r.stderr, r.stderr[-500:]


def write_ply(path, pts):
    with open(path, "wb") as f:
        f.write(("ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\n"
                 "property float z\nproperty float scalar_intensity\nend_header\n" % len(pts)).encode())
        a = np.ascontiguousarray(pts, dtype="<f4").copy()
        a[:, 3] = 0.5
        f.write(a.tobytes())


@pytest.mark.gpu
def test_facade_reference_known_answers_on_gpu(facade_test_exe):
    r = subprocess.run([facade_test_exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-2000:])
    assert " 0 failed" in r.stdout


@pytest.mark.gpu
def test_example_registration_on_bundled_pair(example_exe, bundled, tmp_path):
    """The restated example (GICP / LM / Geman-McClure, robust scale 10 -> 2.5 x3, 1000-point random
    sampling) on the bundled pair's voxelised clouds lands within 10 cm of the data set's own
    T_target_source (the reference ships that file but no test uses it: SURVEY §4)."""
    src, tgt = tmp_path / "source.ply", tmp_path / "target.ply"
    write_ply(src, bundled["source_ds"])
    write_ply(tgt, bundled["target_ds"])
    gt = tmp_path / "T.txt"
    np.savetxt(gt, bundled["T_target_source"])
    r = subprocess.run([example_exe, str(src), str(tgt), "3", str(gt)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    line = [l for l in r.stdout.splitlines() if l.startswith("translation error vs ground truth")][0]
    assert float(line.split(":")[1].split()[0]) < 0.10, r.stdout
    assert "7. Registration" in r.stdout


REF_TESTS = ["test_kdtree", "test_registration_pipeline", "test_downsampling_filters", "test_voxel_hash_map", "test_relative_pose_deskew", "test_preprocess_filter", "test_octree"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", REF_TESTS)
def test_reference_gtest_source_passes_on_libspx(name, bundled, tmp_path):
    """The reference's own cpp/tests/<name>.cpp, compiled UNMODIFIED against include/ + libspx.so through
    tests/cpp/gtest_shim (built by __graft_entry__.build() where /root/reference exists; the binary travels to
    the GPU box): every reference-held assertion runs on the CUDA path."""
    exe = os.path.join(BUILD, "ref_" + name)
    if not os.path.exists(exe):
        pytest.skip("prebuilt reference test binary is absent (no /root/reference at build time)")
    # test_octree.cpp looks for data/target.ply and data/source.ply relative to its working directory (the reference's
    # bundled scans; /root/reference does not exist on the GPU box): it gets the committed fixture clouds there
    (tmp_path / "data").mkdir()
    write_ply(tmp_path / "data" / "target.ply", bundled["target_ds"])
    write_ply(tmp_path / "data" / "source.ply", bundled["source_ds"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900, cwd=str(tmp_path))
    assert r.returncode == 0, (r.stdout[-4000:], r.stderr[-2000:])
    assert " 0 failed" in r.stdout and "[  FAILED  ]" not in r.stdout


@pytest.fixture(scope="module")
def odometry_exe():
    return compile_cpp("examples/example_lidar_odometry.cpp", "example_lidar_odometry")


def test_odometry_example_compiles(odometry_exe):
    assert os.access(odometry_exe, os.X_OK)


@pytest.mark.gpu
def test_cpp_lidar_odometry_pipeline_tracks_the_drive_like_the_python_mirror(odometry_exe, tmp_path):
    """pipeline::lidar_odometry::LiDAROdometryPipeline (include/sycl_points/pipeline/*.hpp; reference:
    pipeline/lidar_odometry.hpp:115-298) on the synthetic drive of tests/test_gpu_odometry.py: every pose within 5 cm of
    the ground truth and within 2 mm of what the Python mirror computes on the same scans (same kernels, same RNG
    streams; the two hosts differ in the rounding of the motion prediction and in the order of the submap's atomics)."""
    import sycl_points_b200 as spx
    from sycl_points_b200 import pipeline as pl
    from test_gpu_odometry import drive, make_params, pose_err
    n = 8
    poses, scans = drive(n)
    for k, s in enumerate(scans):
        write_ply(tmp_path / f"scan_{k:03d}.ply", s)
    np.savetxt(tmp_path / "pose0.txt", poses[0])
    r = subprocess.run([odometry_exe, str(tmp_path), str(n), "0.1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    got = [np.array(l.split("pose")[1].split(), np.float64).reshape(4, 4) for l in r.stdout.splitlines() if l.startswith("frame")]
    assert len(got) == n
    q = spx.DeviceQueue(0)
    P = make_params(pl, spx)
    P.initial_pose = poses[0]
    pipe = pl.LiDAROdometryPipeline(P, q)
    for k in range(n):
        pipe.process(spx.PointCloudShared(q, scans[k]), 0.1 * k)
        dt, da = pose_err(poses[k], got[k].astype(np.float32))
        assert dt < 0.05 and da < 0.005, f"frame {k}: C++ pose {dt:.3f} m / {da:.4f} rad off the ground truth"
        dt2, da2 = pose_err(pipe.get_odom(), got[k].astype(np.float32))
        assert dt2 < 2e-3 and da2 < 2e-4, f"frame {k}: C++ vs Python mirror {dt2:.2e} m / {da2:.2e} rad"
    assert "keyframes" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["example_registration", "example_point_cloud"])
def test_reference_example_source_runs_on_libspx(name, bundled, tmp_path):
    """The reference's own cpp/examples/<name>.cpp, compiled UNMODIFIED against include/ + libspx.so (binary built where
    /root/reference exists), run from a build-like directory next to data/source.ply and data/target.ply."""
    exe = os.path.join(BUILD, "ref_" + name)
    if not os.path.exists(exe):
        pytest.skip("prebuilt reference example binary is absent (no /root/reference at build time)")
    (tmp_path / "data").mkdir()
    (tmp_path / "build").mkdir()
    write_ply(tmp_path / "data" / "target.ply", bundled["target_ds"])
    write_ply(tmp_path / "data" / "source.ply", bundled["source_ds"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900, cwd=str(tmp_path / "build"))
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-2000:])
    if name == "example_registration":
        assert "Registration" in r.stdout, r.stdout[-2000:]
    else:
        assert "Voxel downsampling" in r.stdout and "Compute covariances" in r.stdout, r.stdout[-2000:]
