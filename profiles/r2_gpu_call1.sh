set -x
mkdir -p gpurun_out
python tests/diag_tc_rounding.py > gpurun_out/s1_rounding.txt 2>&1
python -m pytest tests/test_gpu_ops.py -q -x -k "gemm" 2>&1 | tail -15 > gpurun_out/s1_ops.log
python -m pytest tests/test_gpu_model.py tests/test_gpu_zz_reference_graph.py -q 2>&1 | tail -40 > gpurun_out/s1_model.log
MODE=tf32 python tests/diag_gemm_shapes.py > gpurun_out/s1_shapes_tf32.txt 2>&1
MODE=tf32x3 python tests/diag_gemm_shapes.py > gpurun_out/s1_shapes_x3.txt 2>&1
python bench.py --gemm tf32x3 > gpurun_out/s1_bench_x3.json 2> gpurun_out/s1_bench_x3.err
