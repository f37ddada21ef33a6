set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s27_pytest_gpu.txt 2>&1
timeout 300 python bench.py > gpurun_out/s27_bench.json 2> gpurun_out/s27_bench.err
timeout 300 python bench.py --workload st > gpurun_out/s27_bench_st.json 2> gpurun_out/s27_bench_st.err
timeout 200 python examples/train_synthetic.py --seed 3 --iters 25000 --every 1000 --gemm tf32x3 --log gpurun_out/s27_conv_seed3.log > /dev/null 2>&1
