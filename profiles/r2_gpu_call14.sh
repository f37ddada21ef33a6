set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -15 > gpurun_out/s14_tests.log
timeout 600 python bench.py > gpurun_out/s14_bench.json 2> gpurun_out/s14_bench.err
MODE=tf32 timeout 300 python tests/diag_gemm_shapes.py > gpurun_out/s14_shapes_tf32_auto.txt 2>&1
MODE=tf32x3 timeout 300 python tests/diag_gemm_shapes.py > gpurun_out/s14_shapes_tf32x3_auto.txt 2>&1
