set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s26_pytest_gpu.txt 2>&1
timeout 300 python bench.py > gpurun_out/s26_bench.json 2> gpurun_out/s26_bench.err
timeout 300 python bench.py --workload st > gpurun_out/s26_bench_st.json 2> gpurun_out/s26_bench_st.err
STEPS=2 timeout 200 ncu --set full --clock-control none --import-source on -k regex:st_wb_bwd_ref -s 1 -c 1 -o gpurun_out/s26_st_wb_bwd_ref python tests/diag_train_steps.py > gpurun_out/s26_ncu_ref.log 2>&1
