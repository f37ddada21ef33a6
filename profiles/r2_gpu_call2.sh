set -x
mkdir -p gpurun_out
for P in 256 128; do
  AIR_TC_PAIR=$P timeout 300 python -m pytest tests/test_gpu_ops.py -q -x -k "gemm" 2>&1 | tail -12 > gpurun_out/s2_ops_pair$P.log
done
for M in tf32 tf32x3; do for P in 256 128; do
  AIR_TC_PAIR=$P MODE=$M timeout 300 python tests/diag_gemm_shapes.py > gpurun_out/s2_shapes_${M}_pair$P.txt 2>&1
done; done
AIR_TC_PAIR=0 MODE=tf32x3 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -s 4 -c 1 -o gpurun_out/s2_x3_big python tests/diag_roofline_gemm.py > gpurun_out/s2_ncu_x3.log 2>&1
AIR_TC_PAIR=0 AIR_TC_STAGES=2 MODE=tf32x3 timeout 120 python tests/diag_roofline_gemm.py > gpurun_out/s2_x3_stages2.txt 2>&1
AIR_TC_PAIR=0 AIR_TC_BN=64 MODE=tf32x3 timeout 120 python tests/diag_roofline_gemm.py > gpurun_out/s2_x3_bn64.txt 2>&1
