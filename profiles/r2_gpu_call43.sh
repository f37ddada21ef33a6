set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -k "adam" > gpurun_out/s43_pytest_adam.txt 2>&1
timeout 100 python examples/train_synthetic.py --seed 7 --skip-nonfinite --iters 25000 --every 5000 --gemm tf32x3 --log gpurun_out/s43_conv_seed7_skip.log > /dev/null 2>&1
timeout 300 python -m pytest tests/test_gpu_gemm_variants.py -q > gpurun_out/s43_pytest_variants.txt 2>&1
