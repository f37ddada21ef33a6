set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_st.py tests/test_gpu_baseline_sizes.py tests/test_gpu_model.py -q -x > gpurun_out/s50_pytest.txt 2>&1
timeout 300 python bench.py --workload st > gpurun_out/s50_bench_st.json 2> gpurun_out/s50_bench_st.err
timeout 300 python bench.py --no-convergence > gpurun_out/s50_bench_train.json 2> gpurun_out/s50_bench_train.err
