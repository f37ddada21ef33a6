set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s31_pytest_gpu.txt 2>&1
timeout 400 python bench.py > gpurun_out/s31_bench.json 2> gpurun_out/s31_bench.err
timeout 300 python bench.py --workload st > gpurun_out/s31_bench_st.json 2> gpurun_out/s31_bench_st.err
M_FULL=1 timeout 300 python tests/diag_gemm_shapes.py > gpurun_out/s31_gemm_shapes_auto.txt 2>&1
