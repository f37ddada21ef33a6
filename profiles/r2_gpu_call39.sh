set -x
mkdir -p gpurun_out
AIR_TC_CLUSTER4=1 SHAPE="fwd r1 softplus" timeout 60 python tests/diag_one_gemm.py > gpurun_out/s39_one.txt 2>&1
echo "rc=$?" >> gpurun_out/s39_one.txt
AIR_TC_CLUSTER4=1 timeout 300 python -m pytest tests/test_gpu_ops.py -q -x -k "gemm" > gpurun_out/s39_pytest_gemm_cluster4.txt 2>&1
AIR_TC_CLUSTER4=1 M_FULL=1 timeout 200 python tests/diag_gemm_shapes.py > gpurun_out/s39_gemm_shapes_cluster4.txt 2>&1
