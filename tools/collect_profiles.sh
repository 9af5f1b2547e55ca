set -x
python bench.py > gpurun_out/r04a_bench.json 2> gpurun_out/r04a_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r04a_bench_reference_arm.json 2>> gpurun_out/r04a_bench.err
python tools/run_kernels.py all > gpurun_out/r04a_kernels.txt 2>&1
python tools/phase_times.py > gpurun_out/r04a_phase_times.txt 2>&1
python tools/step_timeline.py > gpurun_out/r04a_step_timeline.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r04a_launches_bench.csv python bench.py --steps 3 --warmup 3 > gpurun_out/r04a_ncu_bench.log 2>&1
python profiles/summarize_launches.py gpurun_out/r04a_launches_bench.csv 6 > gpurun_out/r04a_launches_bench_summary.txt
ncu --set full --clock-control none --import-source on -k regex:"align_gn_kernel|voxel_reduce|onesweep|key_hist" -c 12 -o gpurun_out/r04a_full python bench.py --steps 1 --warmup 3 > gpurun_out/r04a_ncu_full.log 2>&1
ls -la gpurun_out/ | tail -15
