set -x
mkdir -p gpurun_out
AIR_TC_PERSIST=0 AIR_TC_PAIR=256 M_FULL=1 timeout 300 python tests/diag_gemm_shapes.py > gpurun_out/s30_gemm_shapes_pair256.txt 2>&1
AIR_TC_PERSIST=0 AIR_TC_PAIR=128 M_FULL=1 timeout 300 python tests/diag_gemm_shapes.py > gpurun_out/s30_gemm_shapes_pair128.txt 2>&1
AIR_TC_PERSIST=0 AIR_TC_PAIR=0 AIR_TC_BN=256 M_FULL=1 timeout 300 python tests/diag_gemm_shapes.py > gpurun_out/s30_gemm_shapes_bn256.txt 2>&1
AIR_TC_PERSIST=1 AIR_TC_PAIR=0 AIR_TC_BN=64 M_FULL=1 timeout 300 python tests/diag_gemm_shapes.py > gpurun_out/s30_gemm_shapes_bn64_persist1.txt 2>&1
AIR_TC_PERSIST=0 AIR_TC_PAIR=0 AIR_TC_BN=64 M_FULL=1 timeout 300 python tests/diag_gemm_shapes.py > gpurun_out/s30_gemm_shapes_bn64.txt 2>&1
