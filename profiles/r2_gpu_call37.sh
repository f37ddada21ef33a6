set -x
mkdir -p gpurun_out
for W in 1 2 3 4; do
  AIR_TC_SPLIT_WAVES=$W M_FULL=1 timeout 300 python tests/diag_gemm_shapes.py 2>&1 | grep -E "^dW|total" > gpurun_out/s37_dw_waves$W.txt
done
