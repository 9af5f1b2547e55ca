T=${1:-r2n}
ncu --set full --clock-control none --profile-from-start off -o /tmp/${T}_bf python tools/bf_experiment.py prof > gpurun_out/${T}_ncu_bf.log 2>&1
ncu -i /tmp/${T}_bf.ncu-rep --page raw --csv > /tmp/${T}_bf.csv
python profiles/ncu_metrics.py /tmp/${T}_bf.csv > gpurun_out/${T}_ncu_bf_p2b2.txt
python - <<PY >> gpurun_out/${T}_ncu_bf_p2b2.txt
import csv
rows=list(csv.reader(open('/tmp/${T}_bf.csv')))
h,u,r=rows[0],rows[1],rows[2]
for i,n in enumerate(h):
    if ('inst_executed_pipe' in n and 'pct' in n) or ('pipe_fma' in n) or ('pipe_alu' in n and 'pct' in n) or 'issue_active' in n:
        print(f"{n:100s} {r[i]:>18s} {u[i]}")
PY
grep -v "^---" gpurun_out/${T}_ncu_bf_p2b2.txt | cut -c1-140 | head -60
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
python -m pytest tests/test_gpu_knn.py tests/test_gpu_registration.py -q -m gpu -k "handles or intensity or bruteforce" 2>&1 | tail -n 3
