set -x
mkdir -p gpurun_out
for M in tf32 tf32x3; do for P in 256 128; do
  AIR_TC_PAIR=$P MODE=$M timeout 300 python tests/diag_gemm_shapes.py > gpurun_out/s8_shapes_${M}_pair$P.txt 2>&1
done; done
AIR_TC_PAIR=0 MODE=tf32 timeout 300 python tests/diag_gemm_shapes.py > gpurun_out/s8_shapes_tf32_1sm.txt 2>&1
AIR_TC_PAIR=0 MODE=tf32x3 timeout 300 python tests/diag_gemm_shapes.py > gpurun_out/s8_shapes_tf32x3_1sm.txt 2>&1
AIR_TC_PAIR=0 AIR_TC_CHAINS=3 MODE=tf32x3 timeout 300 python tests/diag_gemm_shapes.py > gpurun_out/s8_shapes_tf32x3_1sm_ch3.txt 2>&1
