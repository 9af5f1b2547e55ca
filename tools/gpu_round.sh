# one gpurun call: diagnostics + new GPU tests + the whole GPU suite — everything into gpurun_out/<tag>_*
T=${1:-r2c}
python tools/voxelmap_diag.py > gpurun_out/${T}_voxelmap_diag.txt 2>&1; echo "diag rc=$?"
tests/cpp/_build/ref_test_voxel_hash_map > gpurun_out/${T}_ref_voxel_hash_map.txt 2>&1; echo "ref gtest rc=$?"
python -m pytest tests/test_gpu_voxelmap.py -q -m gpu > gpurun_out/${T}_voxelmap_tests.log 2>&1; echo "voxelmap rc=$?"
python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_voxelmap.py::test_lidar_sequence_matches_oracle > gpurun_out/${T}_tests.log 2>&1; echo "suite rc=$?"
tail -n 3 gpurun_out/${T}_voxelmap_tests.log gpurun_out/${T}_tests.log gpurun_out/${T}_ref_voxel_hash_map.txt
