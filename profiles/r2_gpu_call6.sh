set -x
mkdir -p gpurun_out
AIR_TC_PAIR=256 timeout 300 python -m pytest tests/test_gpu_ops.py -q -x -k "gemm or canvas_generator" 2>&1 | tail -12 > gpurun_out/s6_ops_pair256.log
AIR_TC_PAIR=256 MODE=tf32 timeout 300 python tests/diag_gemm_shapes.py > gpurun_out/s6_shapes_tf32_pair256.txt 2>&1
AIR_TC_PAIR=128 MODE=tf32 timeout 300 python tests/diag_gemm_shapes.py > gpurun_out/s6_shapes_tf32_pair128.txt 2>&1
AIR_TC_PAIR=256 AIR_TC_FLAGS=1 MODE=tf32x3 timeout 300 python tests/diag_gemm_shapes.py > gpurun_out/s6_shapes_tf32x3_pair256_f1.txt 2>&1
AIR_TC_PAIR=256 MODE=tf32 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32_pair -s 4 -c 1 -o gpurun_out/s6_pair256_big python tests/diag_roofline_gemm.py > gpurun_out/s6_ncu.log 2>&1
