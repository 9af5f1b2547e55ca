# One gpurun call of round-2 evidence: plain runs first, then ONE ncu pass (the launch list of the same bench command).
# usage (on the GPU box): bash tools/collect_profiles_r2.sh <tag>     -> gpurun_out/<tag>_*
set -x
T=${1:-r2}
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference_arm.json 2>> gpurun_out/${T}_bench.err
python tools/run_kernels.py all > gpurun_out/${T}_kernels.txt 2>&1
python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${T}_bench_short.json 2>> gpurun_out/${T}_bench.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${T}_ncu_bench.log 2>&1
python profiles/summarize_launches.py gpurun_out/${T}_launches_bench.csv 6 > gpurun_out/${T}_launches_bench_summary.txt
ls -la gpurun_out/ | tail -15
