set -x
mkdir -p gpurun_out
for P in 256; do
  AIR_TC_PAIR=$P timeout 300 python -m pytest tests/test_gpu_ops.py -q -x -k "gemm or canvas_generator" 2>&1 | tail -12 > gpurun_out/s5_ops_pair$P.log
done
for M in tf32 tf32x3; do for P in 256 128; do
  AIR_TC_PAIR=$P MODE=$M timeout 300 python tests/diag_gemm_shapes.py > gpurun_out/s5_shapes_${M}_pair$P.txt 2>&1
done; done
