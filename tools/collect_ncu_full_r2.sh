# One gpurun call: `ncu --set full` of one instance of every hot kernel (tools/prof_targets.py brackets them with
# cudaProfilerStart/Stop), after the same program has exited 0 without ncu.  The report stays in /tmp on the box
# (it is larger than what gpurun_out/ may carry back); the raw page comes back as CSV.
# usage: bash tools/collect_ncu_full_r2.sh <tag>
set -x
T=${1:-r2}
python tools/numa_diag.py > gpurun_out/${T}_numa_diag.txt 2>&1
python tools/prof_targets.py 16 > gpurun_out/${T}_prof_targets.txt 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -o /tmp/${T}_full \
    python tools/prof_targets.py 16 > gpurun_out/${T}_ncu_full.log 2>&1
ncu -i /tmp/${T}_full.ncu-rep --page raw --csv > gpurun_out/${T}_ncu_full_raw.csv
python profiles/ncu_metrics.py gpurun_out/${T}_ncu_full_raw.csv > gpurun_out/${T}_ncu_full_summary.txt
ls -la gpurun_out/ /tmp/${T}_full.ncu-rep | tail
