set -x
mkdir -p gpurun_out
ITERS=600 ORACLE_ITERS=200 MODE=tf32x3 timeout 600 python tests/diag_training_vs_oracle.py > gpurun_out/s18_train_vs_oracle.txt 2>&1
