T=${1:-r2e}
python -m pytest tests/test_gpu_voxelmap.py tests/test_gpu_odometry.py -q -m gpu > gpurun_out/${T}_new_tests.log 2>&1; echo "voxelmap+odometry rc=$?"
python -m pytest tests/test_gpu_registration.py -q -m gpu -k "addons or degenerate" > gpurun_out/${T}_addons.log 2>&1; echo "addons rc=$?"
tail -n 40 gpurun_out/${T}_new_tests.log gpurun_out/${T}_addons.log | cut -c1-250
