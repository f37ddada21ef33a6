set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_wb_reference_rounding.py "tests/test_gpu_model.py::test_train_parity_non_default_sizes" -q > gpurun_out/s23_pytest_ref.txt 2>&1
for S in 1 2 3; do
  timeout 200 python examples/train_synthetic.py --seed $S --iters 25000 --every 5000 --gemm tf32x3 --log gpurun_out/s23_conv_sep_seed$S.log > /dev/null 2>&1
  AIR_WB_REF_ATOMICS=1 timeout 200 python examples/train_synthetic.py --seed $S --iters 25000 --every 5000 --gemm tf32x3 --log gpurun_out/s23_conv_atomics_seed$S.log > /dev/null 2>&1
done
