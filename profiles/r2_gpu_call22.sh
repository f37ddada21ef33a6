set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_wb_reference_rounding.py -q -x > gpurun_out/s22_pytest_ref.txt 2>&1
ITERS=41 EVERY=10 MODE=tf32x3 timeout 300 python tests/diag_teacher_forced.py > gpurun_out/s22_teacher_forced_ref.txt 2>&1
timeout 400 python examples/train_synthetic.py --iters 25000 --every 1000 --gemm tf32x3 --log gpurun_out/s22_conv_ref_x3.log > /dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s22_pytest_gpu.txt 2>&1
timeout 300 python bench.py > gpurun_out/s22_bench.json 2> gpurun_out/s22_bench.err
