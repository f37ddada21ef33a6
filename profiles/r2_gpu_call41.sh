set -x
mkdir -p gpurun_out
SEED=7 timeout 300 python tests/diag_nan_hunt.py > gpurun_out/s41_nan_hunt.txt 2>&1
