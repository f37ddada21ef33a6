set -x
mkdir -p gpurun_out
ITERS=301 EVERY=10 MODE=tf32x3 timeout 800 python tests/diag_teacher_forced.py > gpurun_out/s19_teacher_forced.txt 2>&1
