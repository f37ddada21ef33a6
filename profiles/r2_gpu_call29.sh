set -x
mkdir -p gpurun_out
AIR_TC_PERSIST=2 timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -k "gemm" > gpurun_out/s29_pytest_gemm_persist2.txt 2>&1
AIR_TC_PERSIST=1 timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -k "gemm" > gpurun_out/s29_pytest_gemm_persist1.txt 2>&1
for P in 0 1 2; do
  AIR_TC_PERSIST=$P M_FULL=1 timeout 300 python tests/diag_gemm_shapes.py > gpurun_out/s29_gemm_shapes_persist$P.txt 2>&1
done
