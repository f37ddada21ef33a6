set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x 2>&1 | tail -12 > gpurun_out/s9_ops.log
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_zz_reference_graph.py -q 2>&1 | tail -30 > gpurun_out/s9_model.log
MODE=tf32 timeout 300 python tests/diag_gemm_shapes.py > gpurun_out/s9_shapes_tf32_auto.txt 2>&1
MODE=tf32x3 timeout 300 python tests/diag_gemm_shapes.py > gpurun_out/s9_shapes_tf32x3_auto.txt 2>&1
timeout 600 python bench.py > gpurun_out/s9_bench.json 2> gpurun_out/s9_bench.err
MODE=tf32x3 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32_pair -s 4 -c 1 -o gpurun_out/s9_x3_pair_big python tests/diag_roofline_gemm.py > gpurun_out/s9_ncu.log 2>&1
