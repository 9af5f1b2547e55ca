# The C++ facade's odometry loop (examples/example_lidar_odometry.cpp) on the synthetic drive: writes the scans as PLY,
# builds the example against include/ + libspx.so, runs it.  usage (GPU box): bash tools/run_odometry_cpp.sh [frames]
N=${1:-45}
D=/tmp/spx_drive
mkdir -p $D
python - <<PY
import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import synthetic
from test_cpp_facade import write_ply
poses, scans = synthetic.drive($N)
for k, s in enumerate(scans):
    write_ply("$D/scan_%03d.ply" % k, s)
np.savetxt("$D/pose0.txt", poses[0])
np.save("$D/poses.npy", np.array(poses))
PY
g++ -std=c++20 -O2 -Iinclude examples/example_lidar_odometry.cpp -Lsycl_points_b200 -lspx -Wl,-rpath,$PWD/sycl_points_b200 -o /tmp/example_lidar_odometry
/tmp/example_lidar_odometry $D $N 0.1 > /tmp/odo_cpp.txt
python - <<PY
import numpy as np
poses = np.load("$D/poses.npy")
lines = [l for l in open("/tmp/odo_cpp.txt") if l.startswith("frame")]
worst = 0.0
for k, l in enumerate(lines):
    T = np.array(l.split("pose")[1].split(), float).reshape(4, 4)
    worst = max(worst, np.linalg.norm(T[:3, 3] - poses[k][:3, 3]))
print(f"C++ facade LiDAROdometryPipeline, {len(lines)} frames of ~64 k points: worst position error vs ground truth {worst * 100:.1f} cm")
print("".join(l for l in open("/tmp/odo_cpp.txt") if not l.startswith("frame")), end="")
PY
