T=${1:-r2m}
for t in test_relative_pose_deskew test_preprocess_filter; do
  tests/cpp/_build/ref_$t > gpurun_out/${T}_ref_$t.txt 2>&1; echo "$t rc=$?"; tail -n 3 gpurun_out/${T}_ref_$t.txt
done
python -m pytest tests/test_gpu_features_filters.py tests/test_cpp_facade.py tests/test_gpu_odometry.py -q -m gpu 2>&1 | tail -n 8
