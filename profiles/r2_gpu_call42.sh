set -x
mkdir -p gpurun_out
SEED=7 CAPTURE=1 timeout 300 python tests/diag_nan_hunt.py > gpurun_out/s42_nan_hunt_graph.txt 2>&1
SEED=7 CAPTURE=1 WITH_EVAL=1 timeout 300 python tests/diag_nan_hunt.py > gpurun_out/s42_nan_hunt_graph_eval.txt 2>&1
