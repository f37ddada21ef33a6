set -x
mkdir -p gpurun_out
timeout 300 python bench.py --workload st > gpurun_out/s46_bench_st.json 2> gpurun_out/s46_bench_st.err
