T=${1:-r2k}
python tools/bf_experiment.py > gpurun_out/${T}_bf_variants.txt 2>&1
cat gpurun_out/${T}_bf_variants.txt
