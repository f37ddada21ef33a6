set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -8 > gpurun_out/s15_tests.log
timeout 600 python bench.py > gpurun_out/s15_bench.json 2> gpurun_out/s15_bench.err
MODE=tf32 timeout 300 python tests/diag_gemm_shapes.py > gpurun_out/s15_shapes_tf32_auto.txt 2>&1
AIR_TC_BN=256 AIR_TC_PAIR=0 MODE=tf32 timeout 300 python tests/diag_gemm_shapes.py > gpurun_out/s15_shapes_tf32_bn256.txt 2>&1
AIR_TC_BN=256 AIR_TC_PAIR=0 timeout 300 python -m pytest tests/test_gpu_ops.py -q -x -k "gemm" 2>&1 | tail -5 > gpurun_out/s15_ops_bn256.log
