set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_wb_reference_rounding.py -q -k "not training" > gpurun_out/s45_pytest_ref.txt 2>&1
timeout 300 python bench.py --workload st > gpurun_out/s45_bench_st.json 2> gpurun_out/s45_bench_st.err
