set -x
mkdir -p gpurun_out
for P in train+st+infer train+infer train+st; do
  PHASE=$P timeout 300 python tests/diag_capture_after.py > gpurun_out/s34_capture_after_$P.txt 2>&1
done
